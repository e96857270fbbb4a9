"""Top SASS instructions of an `ncu --page source --csv` export by stall samples, plus a per-region instruction count."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
hdr = rows[hi]
end = his[1] - 1 if len(his) > 1 else len(rows)
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[0] != "Address"]
ci = {h: i for i, h in enumerate(hdr)}
f = lambda r, k: float(r[ci[k]] or 0)
tot_inst = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
print("kernel:", rows[0][1][:60], "| total warp-inst", tot_inst, "| samples", tot_samp, "| SASS rows", len(data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    print(f'{r[ci["Address"]][-5:]} samp={f(r,"# Samples"):7.0f} ({100*f(r,"# Samples")/tot_samp:4.1f}%) inst={f(r,"Instructions Executed"):10.0f} thr={r[ci["Avg. Threads Executed"]][:5]:>5}  {r[ci["Source"]][:100]}')
print("--- instruction-count profile (every 16th row) ---")
acc = 0
for j, r in enumerate(data):
    acc += f(r, "Instructions Executed")
    if j % 24 == 23:
        print(f'{r[ci["Address"]][-5:]} cum_inst={acc/tot_inst*100:5.1f}%  last: {r[ci["Source"]][:70]}')
