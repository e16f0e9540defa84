mkdir -p gpurun_out
timeout -s KILL 120 scripts/probes/mma_probe.bin > gpurun_out/r2_mma_probe.txt 2>&1; cat gpurun_out/r2_mma_probe.txt
