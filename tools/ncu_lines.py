#!/usr/bin/env python
"""Per-CUDA-source-line executed warp instructions of one kernel: tools/ncu_lines.py <report> <kernel-regex> [min-share%]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; items = []; funcs = 0
for r in rows:
    if r and r[0] == "Function Name":
        funcs += 1
        if funcs > 1: break
        print(r[1][:110]); continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or r[0] == "": continue
    try: n = int(r[hdr.index("Instructions Executed")])
    except (ValueError, IndexError): continue
    items.append((n, r[0], r[1], r[hdr.index("# Samples")], r[hdr.index("L1 Wavefronts Shared Excessive")]))
tot = sum(i[0] for i in items)
print("total", tot)
for n, line, src, smp, exc in items:
    if tot and 100.0 * n / tot >= thr: print(f"{100.0*n/tot:5.1f}% {n:11d} smp {smp:>6s} exc {exc:>9s} L{line:>5s}: {src.strip()[:110]}")
