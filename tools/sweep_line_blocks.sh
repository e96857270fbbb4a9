#!/bin/bash
# A/B the occupancy target of k_forward_line on the GPU box (rebuilds the library per value).
F=depth-map-fusion-utils_b200/csrc/dmf_forward.cuh
for mb in "$@"; do
  DMF_NVCC_EXTRA="-DDMF_LINE_MIN_BLOCKS=$mb" python depth-map-fusion-utils_b200/build.py --force --verbose 2>&1 | grep -A2 "k_forward_lineILi0ELb1" | tail -1
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('minBlocks=$mb', 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
