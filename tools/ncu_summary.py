#!/usr/bin/env python
"""Summarises .ncu-rep files (ncu -i ... --page raw --csv) into a markdown table: python tools/ncu_summary.py rep1 rep2 ... > profiles/x.md"""
import csv, subprocess, sys
WANT = [("duration", "gpu__time_duration.sum"), ("SM clock", "smsp__cycles_elapsed.avg.per_second"), ("regs/thread", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
        ("block", "launch__block_size"), ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue slots active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("FMA pipe active %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"), ("DRAM throughput %", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("DRAM read", "dram__bytes_read.sum"), ("DRAM written", "dram__bytes_write.sum"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
        ("warp instructions", "smsp__inst_executed.sum")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print(f"## {rep}: empty\n")
        continue
    hdr, units = rows[0], rows[1]
    print(f"## {rep.split('/')[-1]}\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"`{d.get('Kernel Name', '')[:110]}`\n")
        print("| metric | value |\n|---|---|")
        for label, key in WANT:
            if key in d:
                print(f"| {label} | {d[key]} {units[hdr.index(key)]} |")
        st = sorted(((float(d[k]), k) for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k] not in ("", "n/a")), reverse=True)[:5]
        for v, k in st:
            print(f"| stall: {k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')} | {v:.2f} per issue |")
        print()
