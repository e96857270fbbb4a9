mkdir -p gpurun_out
DMF_FWD_CHUNKS=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_forward_line -s 3 -c 1 -o gpurun_out/r01_carve_sign python tools/carve_ab.py S512 64 2 > gpurun_out/r01_carve_sign_ncu.log 2>&1
echo "ncu rc=$?"
timeout 200 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_final2.json
