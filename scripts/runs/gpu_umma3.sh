timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm or one_hot or wqlinear or reference" 2>&1 | tail -3
echo "== mc auto"; timeout -s KILL 300 python scripts/microbench.py --m 256,512,2048 --tag _mc 2>&1 | cut -c1-100
echo "== mc off"; B200AWQ_UMMA_MC=0 timeout -s KILL 300 python scripts/microbench.py --m 2048 --tag _mc0 2>&1 | cut -c1-100
