mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
B200AWQ_FLAT_MIN_M=1 B200AWQ_FLAT_MAX_M=64 timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
echo "== default dispatch with reference"; timeout -s KILL 400 python scripts/microbench.py --m 1,2,4,8,16,32,64,128,256 --ref --tag _sweep 2>&1 | cut -c1-150
echo "== flat for everything <= 64"; B200AWQ_FLAT_MIN_M=1 B200AWQ_FLAT_MAX_M=64 timeout -s KILL 300 python scripts/microbench.py --m 1,32,64 --shapes sq --tag _flatall 2>&1 | cut -c1-100
