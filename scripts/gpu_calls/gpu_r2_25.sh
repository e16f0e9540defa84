# round 2, call 25: L2 wave-ahead prefetch of the CTA's weight slice (B200AWQ_DECODE_L2PF=1 vs 0)
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity_r2.py -q -m gpu -x -k "decode or ring_partition" > gpurun_out/r2_25_tests.log 2>&1; tail -2 gpurun_out/r2_25_tests.log
timeout -s KILL 200 python scripts/microbench.py --m 1 --tag _r25_l2pf 2>&1 | cut -c1-140
B200AWQ_DECODE_L2PF=0 timeout -s KILL 200 python scripts/microbench.py --m 1 --tag _r25_nopf 2>&1 | cut -c1-140
timeout -s KILL 100 python scripts/ring_probe.py 14336 4096 1 48 > gpurun_out/r2_25_probe_14336x4096.txt 2>&1; grep -A3 "^== decode kernel (16" gpurun_out/r2_25_probe_14336x4096.txt | cut -c1-170
