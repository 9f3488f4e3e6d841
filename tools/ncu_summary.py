#!/usr/bin/env python
"""Per-kernel summary table of an ncu report (raw page): python tools/ncu_summary.py rep.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel", 34), ("gpu__time_duration.sum", "ms", 8), ("dram__bytes_read.sum", "rd_GB", 8),
        ("dram__bytes_write.sum", "wr_GB", 8), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 7),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1wf%", 7),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 6),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%", 7),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 7),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 6), ("launch__registers_per_thread", "regs", 5),
        ("launch__grid_size", "grid", 8), ("smsp__inst_executed.sum", "winst", 12),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf", 11),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankcf", 10)]
print(" ".join(f"{c[1]:>{c[2]}s}" for c in cols))
for r in data:
    vals = []
    for name, _, w in cols:
        v = r[ix[name]] if name in ix else "-"
        if name == "Kernel Name":
            v = v.replace("void ", "").replace("(FastStage)", "").replace("p3d::fast::", "")[:w]
        else:
            try:
                f = float(v)
                v = f"{f:.3f}" if f < 1000 else f"{f:.0f}"
            except ValueError:
                pass
        vals.append(f"{v:>{w}s}")
    print(" ".join(vals))
stall = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
for r in data:
    tot = sum(float(r[ix[h]] or 0) for h in stall)
    top = sorted(((h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[ix[h]] or 0)) for h in stall), key=lambda kv: -kv[1])[:7]
    print(r[ix["Kernel Name"]][5:40], " ".join(f"{k}={100 * v / tot:.0f}%" for k, v in top))
