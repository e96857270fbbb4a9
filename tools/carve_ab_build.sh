#!/bin/bash
# Cross-compile carve-kernel variants into build/variants/ (git-ignored, travels to the GPU box); tools/carve_ab_run.sh then
# times each against the in-tree default and checks that the observed grids hash alike.  Flags: csrc/dmf_forward.cuh.
set -e
rm -rf build/variants; mkdir -p build/variants
b() { DMF_B200_OUT=$PWD/build/variants/$1.so DMF_NVCC_EXTRA="$2" python depth-map-fusion-utils_b200/build.py --force --verbose 2>&1 | grep -A2 "k_forward_lineILi0ELb0ELb1" | grep -E "registers|spill" | tr '\n' ' '; echo "built $1 ($2)"; }
b gen1 "-DDMF_CARVE_SIGN=0 -DDMF_CARVE_X2=0 -DDMF_CARVE_MIN_BLOCKS=7" &     # first generation: scalar FP32, per-sample masks
b x2 "-DDMF_CARVE_SIGN=0 -DDMF_CARVE_MIN_BLOCKS=7" &                        # + packed FP32
b signclamp "-DDMF_CARVE_CLAMP=1" &                                         # default with the index clamp
b signb7 "-DDMF_CARVE_MIN_BLOCKS=7" &                                       # default at 72 registers
wait
