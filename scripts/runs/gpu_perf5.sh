mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemv_abi or one_hot" 2>&1 | tail -1
for pad in 0 10240 40960; do
  echo "== pad $pad"
  B200AWQ_STREAM_PAD=$pad timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _v6_pad$pad 2>&1 | cut -c1-100
done
echo "== trace pad 10240"
B200AWQ_STREAM_PAD=10240 timeout -s KILL 300 python scripts/trace_chain.py 4096 4096 1 48 2>&1 | grep -A7 "graph pdl=1" | head -10
echo "== trace pad 40960"
B200AWQ_STREAM_PAD=40960 timeout -s KILL 300 python scripts/trace_chain.py 4096 4096 1 48 2>&1 | grep -A7 "graph pdl=1" | head -10
