#!/usr/bin/env python
"""Condense ncu artefacts from gpurun_out/ into tracked summaries under profiles/.

  python tools/make_profiles.py <tag> <report.ncu-rep> <launches.csv>

writes profiles/<tag>_kernels.txt (per-kernel table + stall mix from the `--set full` capture),
profiles/<tag>_launches.csv (one line per launch: id, kernel, grid, block, ms) and
profiles/traffic.json (DRAM bytes per launch per kernel, read by bench.py for `roofline.traffic`).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep, launches = sys.argv[1:4]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
with open(os.path.join(ROOT, "profiles", f"{tag}_kernels.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none, one launch of each stage kernel of a forward+backward pair\n"
            f"# source report: {os.path.basename(rep)} (not tracked); columns: duration ms, DRAM GB read/written, % of peak\n")
    f.write(summ)

rows = list(csv.reader(open(launches)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[h], rows[h + 1:]
ix = {k: i for i, k in enumerate(hdr)}
tot = {}
with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
    f.write("id,kernel,grid,block,ms\n")
    for r in data:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        if "at::" in name:
            name = "torch::" + name.split("::")[-1][:40]
        ms = float(r[ix["Metric Value"]]) / 1e6
        f.write(f'{r[ix["ID"]]},"{name}","{r[ix["Grid Size"]]}","{r[ix["Block Size"]]}",{ms:.4f}\n')
        tot[name] = tot.get(name, 0.0) + ms
    s = sum(tot.values())
    f.write("# share of all launches in the capture (cold-cache, serialised: compare shares, not absolutes)\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write(f"# {k}: {v:.3f} ms, {100 * v / s:.1f}%\n")

out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
hx = {k: i for i, k in enumerate(rr[0])}
units = rr[1]
traffic = []
for r in rr[2:]:
    def gb(name):
        v, u = float(r[hx[name]]), units[hx[name]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    traffic.append({"kernel": r[hx["Kernel Name"]].replace("void ", "").split("(")[0],
                    "dram_bytes": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
                    "ms": float(r[hx["gpu__time_duration.sum"]])})
order = ["x_r2c", "y_fwd", "z_fwd", "z_bwd", "y_bwd", "x_c2r"]
tj = {"source": os.path.basename(rep), "workload": "1024^3 double, 1x1", "per_launch": {}}
if len(traffic) == 6:
    for k, t in zip(order, traffic):
        tj["per_launch"][k] = t
json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", f"{tag}_kernels.txt")).read())
