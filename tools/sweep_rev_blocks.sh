#!/bin/bash
# A/B the occupancy target of k_reverse on the GPU box (rebuilds the library per value).
F=depth-map-fusion-utils_b200/csrc/dmf_reverse.cuh
for mb in "$@"; do
  sed -i "s/constexpr int REV_MIN_BLOCKS = [0-9]*;/constexpr int REV_MIN_BLOCKS = $mb;/" $F
  python depth-map-fusion-utils_b200/build.py --force --verbose 2>&1 | grep -A2 "k_reverseILb1ELi1" | grep registers | tail -1
  python tools/quick_reverse.py S512 128 2>&1 | grep "fmt=1" | tail -1 | cut -c1-120
done
