for dbg in 63 55 47 31; do
B200AWQ_FUSED_DBG=$dbg timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$dbg scripts/tp_fused_probe.py 2>&1 | grep "N=" | head -2
done
