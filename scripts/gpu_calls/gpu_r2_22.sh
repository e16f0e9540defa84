# round 2, call 22 (2 GPUs): NVLink tensor-parallel tests, then the bench under torchrun with the tp70b object
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout -s KILL 600 python -m pytest tests/test_tp_gpu.py tests/test_attention_gpu.py -q -m gpu > gpurun_out/r2_22_tests.log 2>&1; tail -4 gpurun_out/r2_22_tests.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r2_22_bench_n2.json 2> gpurun_out/r2_22_bench_n2.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_22_bench_n2.json').read().splitlines()[-1]);print(d['metric'],d['value'],d['roofline']['frac']);print(json.dumps(d.get('tp70b'))[:1500])"; tail -3 gpurun_out/r2_22_bench_n2.err
