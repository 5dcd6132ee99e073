#!/usr/bin/env python
"""Per-kernel digest of an `ncu --page raw --csv` export: duration, instructions, IPC, occupancy, DRAM bytes and
the warp-stall breakdown (issue-stalled samples by reason).  usage: ncu_summary.py <raw.csv>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
want = ["gpu__time_duration.sum", "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct"]
stall = [n for n in hdr if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
if not stall:
    stall = [n for n in hdr if "warp_issue_stalled" in n and n.endswith(".pct")]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("==", r[col["Kernel Name"]][:90], "grid", r[col["Grid Size"]] if "Grid Size" in col else "")
    for n in want:
        if n in col:
            print(f"   {n:62s} {r[col[n]]}  {rows[1][col[n]]}")
    ss = sorted(((float(r[col[n]].replace(',', '') or 0), n) for n in stall), reverse=True)[:8]
    for v, n in ss:
        print(f"   stall {n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:.2f}")
