#!/bin/bash
# Build kernel A/B variants: tools/ab_build.sh name "-DFLAG=1 ..." -> build/ab/name.so (development tool)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -cudart static \
  -I include $2 -o build/ab/$1.so metadamage_b200/csrc/mdg_api.cu
