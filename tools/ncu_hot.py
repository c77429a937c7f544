"""Summarise an `ncu --page source --csv` dump: hottest SASS lines with their stall mix."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
data = rows[hi + 1:]
tot = sum(int(r[col["# Samples"]]) for r in data)
agg = {s: sum(int(r[col[s]]) for r in data) for s in stalls}
print("total samples", tot, {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    mix = {s[6:]: int(r[col[s]]) for s in stalls if int(r[col[s]])}
    mix = dict(sorted(mix.items(), key=lambda kv: -kv[1])[:4])
    print(f"{i:5d} {int(r[col['# Samples']]):7d} exec={r[col['Instructions Executed']]:>10s} {r[col['Source']].strip()[:70]:70s} {mix}")
