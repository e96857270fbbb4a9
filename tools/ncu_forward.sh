#!/bin/bash
# ncu evidence for the forward march (run under gpurun, one GPU).  Usage: tools/ncu_forward.sh <tag> [bench args...]
set -u
TAG=${1:-r01}; shift || true
ARGS="--steps 2 --warmup 1 --views 16 --no-cpu-baseline $*"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -c 600 gpurun_out/${TAG}_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_forward -s 1 -c 2 -o gpurun_out/${TAG}_forward python bench.py $ARGS > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
