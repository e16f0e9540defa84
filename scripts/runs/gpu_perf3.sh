mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemv or one_hot or small or engine or wqlinear" 2>&1 | tail -8 > gpurun_out/t3.log
tail -3 gpurun_out/t3.log
timeout -s KILL 300 python scripts/trace_chain.py 4096 4096 1 48 2>&1 | grep -A7 "graph pdl\|plain launches" | head -60
timeout -s KILL 300 python scripts/microbench.py --m 1,2,4,8,16 --tag _v3 > gpurun_out/micro3.log 2>&1; cut -c1-100 gpurun_out/micro3.log
B200AWQ_STREAM_MODE=0 timeout -s KILL 300 python scripts/microbench.py --m 1,4 --tag _v3_mode0 > gpurun_out/micro3_mode0.log 2>&1; cut -c1-100 gpurun_out/micro3_mode0.log
