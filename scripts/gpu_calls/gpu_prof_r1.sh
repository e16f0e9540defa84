mkdir -p gpurun_out
# (1) launch list of the bench command (our kernels only; cold-cache, serialised: compare shares)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:w4a16 -c 1000 --csv --log-file gpurun_out/r1_launches_bench.csv python bench.py --steps 1 --warmup 3 --prefill-steps 3 --no-cpu --no-ref-gpu > gpurun_out/r1_launches_bench.log 2>&1
tail -2 gpurun_out/r1_launches_bench.log | cut -c1-300
# (2) full captures of the two dominant kernels
B200AWQ_PDL=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_stream -s 4 -c 2 -f -o gpurun_out/r1_stream_gemv_14336x4096 python scripts/prof_one.py 1 14336 4096 8 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
B200AWQ_PDL=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_umma -s 2 -c 2 -f -o gpurun_out/r1_umma_gemm_2048x14336x4096 python scripts/prof_one.py 2048 14336 4096 5 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
B200AWQ_PDL=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_flat -s 4 -c 2 -f -o gpurun_out/r1_flat_m16_4096x4096 python scripts/prof_one.py 16 4096 4096 8 > gpurun_out/ncu_c.log 2>&1; tail -1 gpurun_out/ncu_c.log
