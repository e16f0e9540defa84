timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemv or one_hot or gemm_full" 2>&1 | tail -2
timeout -s KILL 300 python scripts/microbench.py --m 1,2,4 --tag _v9 2>&1 | cut -c1-100
timeout -s KILL 300 python scripts/microbench.py --m 1,4 --shapes tp70b --tag _v9 2>&1 | cut -c1-100
