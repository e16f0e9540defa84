# round 2, call 28: TLB warm-up of the activation / output pages before the dependency wait
mkdir -p gpurun_out
timeout -s KILL 200 python scripts/microbench.py --m 1 --tag _r28 2>&1 | cut -c1-140
timeout -s KILL 100 python scripts/ring_probe.py 4096 4096 1 48 > gpurun_out/r2_28_probe_4096x4096.txt 2>&1; grep -A3 "^== decode kernel (16" gpurun_out/r2_28_probe_4096x4096.txt | cut -c1-170
timeout -s KILL 100 python scripts/ring_probe.py 14336 4096 1 48 > gpurun_out/r2_28_probe_14336x4096.txt 2>&1; grep -A3 "^== decode kernel (16" gpurun_out/r2_28_probe_14336x4096.txt | cut -c1-170
