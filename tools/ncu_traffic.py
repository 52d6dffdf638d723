#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of bench.py's default shape (tools/profile.sh):
dram__bytes_read.sum + dram__bytes_write.sum PER FRAME-SET per logical kernel.
usage: tools/ncu_traffic.py <report.ncu-rep> <frame-sets per captured launch> <source tag>"""
import csv, io, json, subprocess, sys

rep, sets, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def num(r, key):
    v = float(r[ix[key]].replace(",", ""))
    u = units[ix[key]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


acc, seen = {}, {}
down = col = 0
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    b = num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    if "cubic5" in name:
        key = "fe_cubic_undistort"; down = 0; col = 0            # a new wave starts
    elif "resize4_walk" in name:
        key = "fe_resize"
    elif "warp_tile" in name:
        key = "warp"; down = 0; col = 0
    elif "pyrdown8_walk" in name:
        key = "pyrdown_l%d" % down; down += 1
    elif "collapse_walk" in name or "collapse8" in name:
        # collapse levels run top-down: the walk + generic launches of one level are adjacent; level 0 kernels are <1>
        lvl = 0 if "<1>" in name else None
        if lvl is None:
            lvl = 2 - (col // 2)
            col += 1
        key = "collapse_l%d" % lvl
    else:
        continue
    acc.setdefault(key, []).append(b)
# collapse levels: sum walk + generic of one wave; others: mean over captured launches
res = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum PER FRAME-SET per kernel from the ncu --set full capture (tools/ncu_traffic.py); "
                   "bench.py multiplies by the frame-sets one launch processes; collapse levels = collapse_walk_kernel + collapse8_kernel of that level",
       "_source": tag, "frame_sets_per_captured_launch": sets}
for k, v in acc.items():
    if k.startswith("collapse"):
        per_wave = sum(v) / max(1, len(v) // 2)
    else:
        per_wave = sum(v) / len(v)
    res[k] = per_wave / sets
print(json.dumps(res, indent=1))
