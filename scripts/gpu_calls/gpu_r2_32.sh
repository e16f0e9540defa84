# round 2, call 32: after removing the superseded int8-digit ring kernel from w4a16_ring.cu: decode-path parity + smoke
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_tp_loopback_gpu.py -q -m gpu -x > gpurun_out/r2_32_tests.log 2>&1; tail -3 gpurun_out/r2_32_tests.log
timeout -s KILL 200 python __graft_entry__.py --smoke 2>&1 | tail -3
B200AWQ_DECODE=0 timeout -s KILL 200 python scripts/microbench.py --m 1,2,4 --shapes sq --tag _r32_ring 2>&1 | cut -c1-140
