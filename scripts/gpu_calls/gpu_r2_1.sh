# round 2, call 1: first run of the persistent ring kernel -- parity, A/B against the round-1 stream kernel, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_1_smi.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemv or one_hot" 2>&1 | tail -40 > gpurun_out/r2_1_t_gemv.log
tail -15 gpurun_out/r2_1_t_gemv.log
B200AWQ_RING=1 timeout -s KILL 300 python scripts/microbench.py --m 1,2,4 --tag _ring 2>&1 | cut -c1-120 | tee gpurun_out/r2_1_mb_ring.log
B200AWQ_RING=0 timeout -s KILL 300 python scripts/microbench.py --m 1,2,4 --tag _stream 2>&1 | cut -c1-120 | tee gpurun_out/r2_1_mb_stream.log
B200AWQ_RING_SPLIT=1 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _ring_s1 2>&1 | cut -c1-120 | tee gpurun_out/r2_1_mb_ring_s1.log
timeout -s KILL 900 python bench.py > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err; cat gpurun_out/r2_1_bench.json | cut -c1-1500; tail -3 gpurun_out/r2_1_bench.err
timeout -s KILL 900 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -5
