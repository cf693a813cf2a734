#!/bin/bash
# builds rtsengine_b200/variants/librts_<name>.so for A/B runs on one box:  tools/build_variants.sh name "-DRTS_X=.. -DRTS_Y=.." [name flags ...]
set -e
cd "$(dirname "$0")/../rtsengine_b200/csrc"
mkdir -p ../variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a $flags -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC -I../../include \
      -ccbin /usr/bin/g++ -shared -o ../variants/librts_$name.so rts_abi.cu &
done
wait
ls -la ../variants
