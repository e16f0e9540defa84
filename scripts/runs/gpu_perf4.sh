mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemv_abi or one_hot" 2>&1 | tail -2
timeout -s KILL 300 python scripts/trace_chain.py 4096 4096 1 48 2>&1 | grep -A7 "graph pdl" | head -40
timeout -s KILL 300 python scripts/trace_chain.py 14336 4096 1 24 2>&1 | grep -A7 "graph pdl=1" | head -20
timeout -s KILL 300 python scripts/microbench.py --m 1,2 --tag _v5 2>&1 | cut -c1-100
B200AWQ_STREAM_RPB=4 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _v5_rpb4 2>&1 | cut -c1-100
