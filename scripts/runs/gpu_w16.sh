timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemv or one_hot" 2>&1 | tail -2
echo "== default (16 warps single-wave, 8 warps + rbs 3 for big N)"; timeout -s KILL 300 python scripts/microbench.py --m 1,2,4 --tag _v7 2>&1 | cut -c1-100
echo "== warps 8 rbs 1 (previous)"; B200AWQ_STREAM_WARPS=8 B200AWQ_STREAM_RBS=1 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _v7_w8 2>&1 | cut -c1-100
echo "== warps 8 rbs 2"; B200AWQ_STREAM_WARPS=8 B200AWQ_STREAM_RBS=2 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _v7_w8r2 2>&1 | cut -c1-100
echo "== warps 16 rbs 3"; B200AWQ_STREAM_WARPS=16 B200AWQ_STREAM_RBS=3 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _v7_w16r3 2>&1 | cut -c1-100
