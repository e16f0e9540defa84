mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_tp_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 scripts/tp_fused_probe.py 2>&1 | grep "N="
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 --workload tp70b > gpurun_out/bench_tp70b_n2.json 2> gpurun_out/bench_tp70b_n2.err; cat gpurun_out/bench_tp70b_n2.json | cut -c1-1600; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/bench_tp70b_n2.err | tail -3
