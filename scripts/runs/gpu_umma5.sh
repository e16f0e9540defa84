timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm or one_hot or reference" 2>&1 | tail -2
echo "== umma2 auto"; timeout -s KILL 300 python scripts/microbench.py --m 64,512,2048 --tag _u3 2>&1 | cut -c1-100
echo "== umma2 off"; B200AWQ_UMMA2=0 timeout -s KILL 300 python scripts/microbench.py --m 512,2048 --tag _u3off 2>&1 | cut -c1-100
