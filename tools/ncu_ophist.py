#!/usr/bin/env python
"""Dynamic SASS opcode histogram of one kernel from an ncu report (source page).
usage: tools/ncu_ophist.py <report.ncu-rep> <kernel-regex> [launch-id]"""
import csv, collections, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat]
if len(sys.argv) > 3: cmd += ["--launch-skip", sys.argv[3], "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; c = collections.Counter(); s = collections.Counter(); tot = 0; kernels = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1: break
        print(r[1][:120]); continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    try: n = int(r[hdr.index("Instructions Executed")])
    except ValueError: continue
    src = r[hdr.index("Source")].split()
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0].rstrip(";")
    c[op] += n; tot += n; s[op] += int(r[hdr.index("# Samples")] or 0)
print("total warp instructions", tot)
for k, v in c.most_common(30): print(f"{k:8s} {v:12d} {100*v/tot:5.1f}%  samples {s[k]}")
