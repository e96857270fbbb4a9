#!/bin/bash
# ncu evidence for the reverse march (run under gpurun, one GPU).  Usage: tools/ncu_reverse.sh <tag>
set -u
TAG=${1:-r01rev}
mkdir -p gpurun_out
python tools/quick_reverse.py S512 64 > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
grep "fmt=1" gpurun_out/${TAG}_plain.log | tail -1
ncu --set full --clock-control none --import-source on -k regex:k_reverse -s 2 -c 1 -o gpurun_out/${TAG}_reverse python tools/quick_reverse.py S512 64 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
