# round 2, call 16: what SM clock does the decode chain actually run at?  (in-kernel cycles / ns, and nvidia-smi at 20 ms)
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,power.limit,clocks_throttle_reasons.active,temperature.gpu --format=csv -lms 20 > gpurun_out/r2_16_smi.csv 2>&1 &
SMI=$!
timeout -s KILL 200 python scripts/ring_probe.py 14336 4096 1 48 > gpurun_out/r2_16_probe_14336x4096.txt 2>&1
timeout -s KILL 200 python scripts/ring_probe.py 4096 4096 1 48 > gpurun_out/r2_16_probe_4096x4096.txt 2>&1
timeout -s KILL 300 python scripts/microbench.py --m 1 --ring-mb 2048 --tag _r16 2>&1 | cut -c1-140
kill $SMI
grep -h "^==\|SM clock" gpurun_out/r2_16_probe_*.txt
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_16_smi.csv')) if len(r)>=5 and r[0].strip().split()[0].isdigit()]
clk=[int(r[0].split()[0]) for r in rows]; pw=[float(r[2].split()[0]) for r in rows]
print("samples",len(rows),"sm clock min/median/max",min(clk),sorted(clk)[len(clk)//2],max(clk),"power max",max(pw), "limit", rows[0][3])
busy=[(c,p,r[4]) for c,p,r in zip(clk,pw,rows) if p>400]
print("samples above 400 W:",len(busy), "clock min/median", (min(b[0] for b in busy), sorted(b[0] for b in busy)[len(busy)//2]) if busy else None, set(b[2].strip() for b in busy))
PY
