timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm or one_hot or reference" 2>&1 | tail -2
timeout -s KILL 300 python scripts/microbench.py --m 64,2048 --tag _u6 2>&1 | cut -c1-100
