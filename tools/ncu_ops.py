"""Opcode histogram (weighted by executed count) of an `ncu --page source --csv` dump."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; col = {n: i for i, n in enumerate(h)}
cnt = collections.Counter()
for r in rows[hi + 1:]:
    src = r[col["Source"]].strip().split()
    op = src[1] if src[0].startswith("@") else src[0]
    cnt[op.split(".")[0]] += int(r[col["Instructions Executed"]])
tot = sum(cnt.values())
print("total", tot)
for k, v in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{k:12s} {v:12d} {100 * v / tot:5.1f}%")
