#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU):  python scripts/ncu_summary.py file.ncu-rep [pattern ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
                        "sm__throughput.avg.pct", "registers_per_thread", "grid_size", "block_size", "occupancy_limit",
                        "waves_per_multiprocessor", "warps_active.avg.pct", "sm__cycles_elapsed.avg ", "sm__cycles_active.avg",
                        "smsp__inst_executed.sum ", "issue_active.avg.pct", "warp_issue_stalled", "shared_mem_per_block",
                        "pipe_tensor", "inst_executed_pipe_alu", "inst_executed_pipe_fma", "bank_conflicts", "lts__t_sector_hit_rate",
                        "achieved_occupancy", "sm__inst_executed.avg.per_cycle_active", "smsp__cycles_active.avg "]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("---", r[hdr.index("Kernel Name")][:70], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for i, h in enumerate(hdr):
        if any(p.strip() in h for p in pats):
            v = r[i]
            if "stalled" in h and "per_warp_active" not in h:
                continue
            try:
                if float(v.replace(",", "")) == 0:
                    continue
            except ValueError:
                pass
            print(f"  {h.split('.', 2)[-1] if h.count('.') > 3 else h:90s} {v} {units[i]}")
