# round 2, call 27: consumer instruction diet (scale and zero terms through one FMA per octet and lane, two groups unrolled)
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_layer_ops_gpu.py -q -m gpu -x -k "gemv or one_hot or decode or kat or mlp_front or ring_partition" > gpurun_out/r2_27_tests.log 2>&1; tail -2 gpurun_out/r2_27_tests.log
timeout -s KILL 200 python scripts/microbench.py --m 1 --tag _r27 2>&1 | cut -c1-140
B200AWQ_DECODE_WARPS=8 timeout -s KILL 200 python scripts/microbench.py --m 1 --tag _r27_w8 2>&1 | cut -c1-140
timeout -s KILL 100 python scripts/ring_probe.py 14336 4096 1 48 > gpurun_out/r2_27_probe_14336x4096.txt 2>&1; grep -A5 "^== decode kernel (16" gpurun_out/r2_27_probe_14336x4096.txt | cut -c1-170; grep "warp 0 " gpurun_out/r2_27_probe_14336x4096.txt | head -2
