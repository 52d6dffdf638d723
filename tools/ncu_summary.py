#!/usr/bin/env python
"""Key per-kernel metrics of an ncu report: tools/ncu_summary.py <report.ncu-rep>"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("l1tex__data_pipe_lsu_wavefronts.sum", "lsu_wf"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("smsp__inst_executed.sum", "inst"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_nsel"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    name = name.split("(")[0].split("::")[-1] + ("<" + name.split("<")[-1].split(">")[0] + ">" if "<" in name.split("(")[0] else "")
    print(name)
    line = []
    for k, short in want:
        if k in idx:
            v = r[idx[k]]
            try: v = float(v.replace(",", "")); v = ("%.0f" % v) if abs(v) >= 100 else ("%.2f" % v)
            except ValueError: pass
            line.append(f"{short}={v}")
    print("   " + "  ".join(line))
