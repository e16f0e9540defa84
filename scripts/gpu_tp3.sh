for dbg in 0 1 2 4 7; do
B200AWQ_FUSED_DBG=$dbg timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$dbg scripts/tp_fused_probe.py 2>&1 | grep "N="
done
