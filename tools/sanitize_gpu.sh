#!/bin/bash
# compute-sanitizer over the smallest end-to-end run of every kernel family (run under gpurun, one GPU; memcheck slows kernels
# 10-100x, so this is smoke() -- forward march, reverse march, carve, id lists on S64 -- plus the API edge tests, not the suite).
#   gpurun --timeout 900 -- 'bash tools/sanitize_gpu.sh'        -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
    timeout 200 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 \
        python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_${tool}_smoke.log 2>&1
    echo "$tool smoke rc=$?"; tail -3 gpurun_out/sanitize_${tool}_smoke.log
done
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 \
    python -m pytest tests/test_api_edge_gpu.py -x -q -m gpu > gpurun_out/sanitize_memcheck_api_edge.log 2>&1
echo "memcheck api_edge rc=$?"; tail -3 gpurun_out/sanitize_memcheck_api_edge.log
