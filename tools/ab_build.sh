#!/bin/bash
# build the library with extra nvcc flags (quoted first arg), then run the remaining command
FLAGS="$1"; shift
DMF_NVCC_EXTRA="$FLAGS" python depth-map-fusion-utils_b200/build.py --force > /dev/null || exit 1
"$@"
