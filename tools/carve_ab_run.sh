#!/bin/bash
# One GPU call: time every carve build under build/variants/ against the in-tree default, keep the fastest one whose observed
# grids (one view, 128 views) hash like the default's, and run the carve parity tests + smoke() on that one.
# Build the variants first (CPU, cross-compiled):  tools/carve_ab_build.sh
mkdir -p gpurun_out
LOG=gpurun_out/carve_ab.log
: > $LOG
python tools/carve_ab.py >> $LOG 2>&1
for so in build/variants/*.so; do
    DMF_B200_LIB=$PWD/$so timeout 120 python tools/carve_ab.py >> $LOG 2>&1
done
cat $LOG
BEST=$(python - <<'PY'
import re
rows = []
for l in open("gpurun_out/carve_ab.log"):
    m = re.match(r"(\S+): .* steady median=([\d.]+) .*sha1\(one view\)=(\w+) sha1\(\d+ views\)=(\w+)", l)
    if m:
        rows.append((m.group(1), float(m.group(2)), m.group(3), m.group(4)))
base = next(r for r in rows if r[0] == "libdmf_b200.so")
ok = [r for r in rows if r[2:] == base[2:] and r[0] != base[0]]
ok.sort(key=lambda r: r[1])
print(ok[0][0] if ok and ok[0][1] < base[1] else "")
PY
)
echo "best variant: '$BEST'" | tee -a $LOG
if [ -n "$BEST" ]; then
    export DMF_B200_LIB=$PWD/build/variants/$BEST
    timeout 300 python -m pytest tests/test_carve_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee -a $LOG
    timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee -a $LOG
fi
