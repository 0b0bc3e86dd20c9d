#!/bin/bash
# Run tools/profile_target.py against every build/ab/*.so (development tool; on the GPU box)
cd "$(dirname "$0")/.."
for so in build/ab/*.so; do
  echo "== $(basename $so .so)"
  MDG_LIB_PATH=$PWD/$so timeout 100 python tools/profile_target.py ${1:-10000} ${2:-500} ${3:-1000} 2>&1 | tail -1
done
