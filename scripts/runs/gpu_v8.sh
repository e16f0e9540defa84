timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemv or one_hot" 2>&1 | tail -2
timeout -s KILL 300 python scripts/microbench.py --m 1,2,4 --tag _v8 2>&1 | cut -c1-100
for pad in 10240 40960 -1; do echo "== pad $pad"; B200AWQ_STREAM_PAD=$pad timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _v8_pad$pad 2>&1 | cut -c1-100; done
