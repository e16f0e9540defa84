mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "one_hot or gemv_abi" 2>&1 | tail -12
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not one_hot and not gemv_abi" 2>&1 | tail -12
timeout -s KILL 300 python scripts/microbench.py --m 1,2,4,8,16,64 --tag _flat1 2>&1 | cut -c1-100
