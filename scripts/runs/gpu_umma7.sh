timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm or one_hot or reference" 2>&1 | tail -2
timeout -s KILL 300 python scripts/microbench.py --m 64,512,2048 --tag _u7 2>&1 | cut -c1-100
echo "== TN 128"; B200AWQ_UMMA_TN=128 timeout -s KILL 300 python scripts/microbench.py --m 2048 --shapes sq --tag _u7tn128 2>&1 | cut -c1-100
