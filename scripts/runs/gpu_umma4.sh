timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm or one_hot or wqlinear or reference" 2>&1 | tail -3
echo "== umma2 auto"; timeout -s KILL 300 python scripts/microbench.py --m 512,1024,2048,4096 --tag _u2 2>&1 | cut -c1-100
echo "== umma2 off"; B200AWQ_UMMA2=0 timeout -s KILL 300 python scripts/microbench.py --m 2048 --tag _u2off 2>&1 | cut -c1-100
