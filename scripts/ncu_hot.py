#!/usr/bin/env python
"""Top stall sites of the first kernel in an .ncu-rep:  python scripts/ncu_hot.py file.ncu-rep [N] [kernel index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40; kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
s0 = starts[kidx]; end = starts[kidx + 1] - 1 if kidx + 1 < len(starts) else len(rows)
h = rows[s0]; body = [r for r in rows[s0 + 1:end] if len(r) > 10]
idx = {n: i for i, n in enumerate(h)}
tot = sum(int(r[idx["# Samples"]]) for r in body)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(int(r[idx[s]]) for r in body) for s in stalls}
print("kernel:", rows[s0 - 1][1][:80] if s0 else "", "| total samples", tot, "| instructions", len(body))
print("stall totals:", [(k, v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v][:10])
# cumulative sample profile along the program (regions of 40 instructions)
cum = 0
print("samples by region of 40 SASS instructions:")
for i in range(0, len(body), 40):
    s = sum(int(r[idx["# Samples"]]) for r in body[i:i + 40])
    ex = max(int(r[idx["Instructions Executed"]]) for r in body[i:i + 40])
    print(f"  [{i:4d}-{i+39:4d}] samples {s:6d} ({100*s/max(tot,1):5.1f}%)  max-exec {ex:7d}  first: {body[i][idx['Source']].strip()[:50]}")
print("top instructions:")
for r in sorted(body, key=lambda r: -int(r[idx["# Samples"]]))[:topn]:
    st = sorted([(s[6:], int(r[idx[s]])) for s in stalls if int(r[idx[s]]) > 0], key=lambda kv: -kv[1])[:3]
    print(r[idx["# Samples"]].rjust(6), r[idx["Instructions Executed"]].rjust(7), str(body.index(r)).rjust(5), r[idx["Source"]].strip()[:64].ljust(64), st)
