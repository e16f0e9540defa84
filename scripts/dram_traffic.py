#!/usr/bin/env python
"""ncu CSV of scripts/dram_step.py -> profiles/traffic.json (DRAM bytes per launch of the decode and the prefill step).
    python scripts/dram_traffic.py gpurun_out/r2_dram_step.csv"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]
per = {}
for r in rows[h + 1:]:
    if len(r) <= H.index("Metric Value"):
        continue
    v = float(r[H.index("Metric Value")].replace(",", ""))
    u = r[H.index("Metric Unit")]
    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    per.setdefault(int(r[0]), [r[H.index("Kernel Name")], 0.0])[1] += v
ids = sorted(per)
dec = [per[i][1] for i in ids if "decode" in per[i][0] or "ring" in per[i][0] or "stream" in per[i][0]]
pre = [per[i][1] for i in ids if "umma" in per[i][0] or "flat" in per[i][0]]
alg_dec = 3712483328.0
out = {"decode_dram_bytes_per_launch": sum(dec) / max(len(dec), 1), "decode_launches_profiled": len(dec),
       "decode_dram_bytes_per_step": sum(dec), "algorithmic_bytes_per_step": alg_dec,
       "prefill_dram_bytes_per_launch": sum(pre) / max(len(pre), 1), "prefill_launches_profiled": len(pre),
       "prefill_dram_bytes_per_step": sum(pre),
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the launches of one decode step and one prefill "
                 "step (scripts/dram_step.py; profiles/r2_dram_step.csv)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
