#!/usr/bin/env python
"""Digest of a tools/profile.sh run into profiles/: launch list, per-kernel shares of one wave next to bench.py's
CUDA-event shares, ncu --set full raw page + key metrics, DRAM traffic per frame-set.
usage: tools/profile_digest.py <tag>   (reads gpurun_out/<tag>_*, writes profiles/<tag>_* and profiles/traffic.json)"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
OURS = ("cubic", "resize4", "warp_tile", "pyrdown8", "pyrdown_kernel", "coarsest", "collapse", "yuyv", "direct_blend", "warp_kernel", "mixed")

shutil.copy(os.path.join(G, tag + "_launches.csv"), os.path.join(P, tag + "_ncu_launches.csv"))
txt = open(os.path.join(G, tag + "_launches.csv")).read()
rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
ours = [r for r in rows if any(k in r["Kernel Name"] for k in OURS)]
bench = json.load(open(os.path.join(G, tag + "_plain.json")))
sets = bench["config"]["frame_sets_per_wave"]
z = str(4 * sets) + ")"
idx = [i for i, r in enumerate(ours) if "cubic5" in r["Kernel Name"] and r["Grid Size"].endswith(", " + z)]
i0, i1 = idx[-2], idx[-1]
tot = sum(float(r["Metric Value"].replace(",", "")) for r in ours[i0:i1])
with open(os.path.join(P, tag + "_launch_shares.txt"), "w") as f:
    f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none) of one wave of %d frame-sets, bench.py default workload (config 2)\n" % sets)
    f.write("# command: python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1   [profiles/%s_ncu_launches.csv]\n" % tag)
    for r in ours[i0:i1]:
        n = r["Kernel Name"].split("(")[0].split("::")[-1]
        us = float(r["Metric Value"].replace(",", "")) / 1000.0
        f.write("%-30s grid %-18s %8.1f us  %5.1f%% of wave\n" % (n, r["Grid Size"], us, 100 * us * 1000 / tot))
    f.write("wave total %.1f us under ncu (cold cache, serialised)\n\n" % (tot / 1000))
    f.write("# bench.py (CUDA events, no profiler) shares of the same command:\n")
    for k, v in bench["roofline"]["kernels"].items():
        f.write("%-22s %8.1f us  %5.1f%%\n" % (k, 1000 * v["ms_per_launch"], 100 * v["share"]))

rep = os.path.join(G, tag + "_full.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(os.path.join(P, tag + "_ncu_full_raw.csv"), "w").write(raw)
summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
open(os.path.join(P, tag + "_ncu_summary.txt"), "w").write(summ)

r = list(csv.reader(io.StringIO(raw)))
hdr = r[0]
kn, gs = hdr.index("Kernel Name"), hdr.index("launch__grid_size")
rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
seen = {}
for row in r[2:]:
    name = row[kn].split("(")[0].split("::")[-1]
    seen.setdefault((name, int(row[gs])), float(row[rd]) * unit[r[1][rd]] + float(row[wr]) * unit[r[1][wr]])
by_kernel = {}
for (name, g), b in seen.items():
    by_kernel.setdefault(name.split("<")[0] + ("<1>" if name.endswith("<1>") else ""), []).append((g, b))
for v in by_kernel.values():
    v.sort(reverse=True)            # largest grid = finest level
def lvl(name, i):
    v = by_kernel.get(name, [])
    return v[i][1] if i < len(v) else 0.0
# DRAM traffic per frame-set and logical kernel: one implementation, tools/ncu_traffic.py (it follows the launch ORDER of a
# wave, which is what tells the collapse levels apart)
t = json.loads(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), os.path.join(G, tag + "_full.ncu-rep"),
                               str(sets), "profiles/%s_ncu_full_raw.csv" % tag], capture_output=True, text=True, check=True).stdout)
json.dump(t, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(open(os.path.join(P, tag + "_launch_shares.txt")).read())
print(json.dumps(t, indent=1))
