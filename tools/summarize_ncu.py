#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a small markdown/JSON table for profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_ncu_full.md [--traffic profiles/traffic.json [--clips N]]
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe (LSU + tensor operand reads) % of peak"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tensor-core operand reads from smem % of peak"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == "--traffic" else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    lines = [f"# ncu --set full summary of `{rep}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200;",
             "per-launch values (cold-ish caches, kernels serialised by the profiler).", ""]
    traffic = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        lines.append(f"## `{name[:110]}`")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k, label in KEYS:
            if k in col:
                lines.append(f"| {label} (`{k}`) | {r[col[k]]} | {units[col[k]]} |")
        lines.append("")
        try:
            def to_bytes(v, u):
                m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                return float(v) * m.get(u, 1)
            rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
            wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            m = re.search(r"(\w+_kernel)", name)
            short = m.group(1) if m else name.split("<")[0].split("::")[-1].replace("void ", "").strip()
            traffic[short] = rd + wr
        except Exception:
            pass
    open(out, "w").write("\n".join(lines) + "\n")
    if traffic_path:
        if len(sys.argv) > 6 and sys.argv[5] == "--clips":
            traffic["clips_per_launch"] = int(sys.argv[6])
        json.dump(traffic, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
