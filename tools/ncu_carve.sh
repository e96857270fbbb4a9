#!/bin/bash
# ncu evidence for the carve kernel (run under gpurun, one GPU): parity tests first, then one --set full capture of a
# steady-state 64-view launch (DMF_FWD_CHUNKS=1: the host API marches the batch as one launch).  Usage: tools/ncu_carve.sh <tag>
TAG=${1:-r01_carve}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_carve_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${TAG}_tests.log
DMF_FWD_CHUNKS=1 python tools/carve_ab.py S512 64 2 > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
cat gpurun_out/${TAG}_plain.log
DMF_FWD_CHUNKS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forward_line -s 3 -c 1 -o gpurun_out/${TAG} python tools/carve_ab.py S512 64 2 > gpurun_out/${TAG}_ncu.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log
ls -la gpurun_out | tail -6
