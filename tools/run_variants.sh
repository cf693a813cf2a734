#!/bin/bash
# A/B of the build variants under rtsengine_b200/variants/ (diagnostic): whole-frame time per variant, same frames
for so in rtsengine_b200/variants/librts_*.so; do
  echo "== $so"
  RTS_B200_LIB=$PWD/$so timeout 300 python tools/ab_flags.py 1000000 only-default 2>&1 | tail -2
done
