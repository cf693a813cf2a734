#!/bin/bash
# A/B of the build variants under rtsengine_b200/variants/ (diagnostic): tools/ab_modes.py per variant, same frames
MODES=${1:-1}
for so in rtsengine_b200/variants/librts_*.so; do
  RTS_B200_LIB=$PWD/$so timeout 300 python tools/ab_modes.py 1000000 $MODES 2>&1 | tail -2
done
