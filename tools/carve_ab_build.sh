#!/bin/bash
# Cross-compile the carve-kernel variants timed by tools/carve_ab_run.sh into build/variants/ (git-ignored, travels to the GPU box).
set -e
rm -rf build/variants; mkdir -p build/variants
b() { DMF_B200_OUT=$PWD/build/variants/$1.so DMF_NVCC_EXTRA="$2" python depth-map-fusion-utils_b200/build.py --force --verbose 2>&1 | grep -A2 "k_forward_lineILi0ELb0ELb1" | grep -E "registers|spill" | tr '\n' ' '; echo "built $1 ($2)"; }
b sign "-DDMF_CARVE_SIGN=1" &
b signb6 "-DDMF_CARVE_SIGN=1 -DDMF_CARVE_MIN_BLOCKS=6" &
b signb8 "-DDMF_CARVE_SIGN=1 -DDMF_CARVE_MIN_BLOCKS=8" &
b signnc "-DDMF_CARVE_SIGN=1 -DDMF_CARVE_CLAMP=0" &
wait
b signncb8 "-DDMF_CARVE_SIGN=1 -DDMF_CARVE_CLAMP=0 -DDMF_CARVE_MIN_BLOCKS=8" &
b signm16b6 "-DDMF_CARVE_SIGN=1 -DDMF_CARVE_MLP=16 -DDMF_CARVE_MIN_BLOCKS=6" &
wait
