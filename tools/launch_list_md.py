#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a small markdown table for profiles/.

    python tools/launch_list_md.py gpurun_out/x.csv profiles/x.md "title / command"
"""
import collections
import csv
import sys


def main():
    src, out, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    col = {n: i for i, n in enumerate(h)}
    agg = collections.OrderedDict()
    n = 0
    for r in rows[hi + 1:]:
        if len(r) < len(h) or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v, u = float(r[col["Metric Value"]]), r[col["Metric Unit"]]
        ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        name = r[col["Kernel Name"]].replace("void ", "").replace("dspb200::", "").replace("<unnamed>::", "")
        name = name.split("(")[0]
        key = (name, r[col["Grid Size"]], r[col["Block Size"]])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ms
        n += 1
    tot = sum(a[1] for a in agg.values())
    lines = [f"# ncu launch list: {title}", "",
             f"`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` (raw rows: `{src.split('/')[-1]}`), "
             f"{n} launches, {tot:.1f} ms of kernel time.",
             "Times under ncu are serialised and cold (the GPU idles between launches, so the SM clock sits near its maximum "
             "instead of the power-capped clock of the back-to-back bench): the SHARES are what compares with the bench line.",
             "", "| kernel | grid | block | launches | avg ms | total ms | share |", "|---|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k[0]}` | {k[1]} | {k[2]} | {a[0]} | {a[1] / a[0]:.3f} | {a[1]:.2f} | {100 * a[1] / tot:.1f} % |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
