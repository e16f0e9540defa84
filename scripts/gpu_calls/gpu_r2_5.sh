# round 2, call 5: per-instruction throughput of HMMA / IMMA / ALU ops on one SM; asymptotic streaming rate of the ring
mkdir -p gpurun_out
timeout -s KILL 120 scripts/probes/mma_probe.bin > gpurun_out/r2_mma_probe.txt 2>&1; cat gpurun_out/r2_mma_probe.txt
timeout -s KILL 300 python scripts/ring_probe.py 114688 4096 1 6 > gpurun_out/r2_5_probe_big.txt 2>&1; grep "^==" gpurun_out/r2_5_probe_big.txt
