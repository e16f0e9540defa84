# round 2, call 23: multi-token variant of the decode kernel (2..4 fp16 tokens): parity, microbench against the ring kernel
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_attention_gpu.py -q -m gpu -x -k "gemv or one_hot or ring or decode or kat or attention or rope or llama2" > gpurun_out/r2_23_tests.log 2>&1; tail -4 gpurun_out/r2_23_tests.log
timeout -s KILL 300 python scripts/microbench.py --m 2,3,4 --tag _r23_mt 2>&1 | cut -c1-140
B200AWQ_DECODE_MT_MAX_M=0 timeout -s KILL 300 python scripts/microbench.py --m 2,4 --tag _r23_ring 2>&1 | cut -c1-140
timeout -s KILL 200 python scripts/ring_probe.py 14336 4096 4 48 > gpurun_out/r2_23_probe_m4.txt 2>&1; grep -A4 "^==" gpurun_out/r2_23_probe_m4.txt | cut -c1-170 | head -30
