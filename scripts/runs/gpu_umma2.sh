timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
timeout -s KILL 400 python scripts/microbench.py --m 17,32,64,128,256,512,2048 --tag _umma2 2>&1 | cut -c1-100
