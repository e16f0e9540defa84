# round 2, call 31 (4 GPUs): the bench under torchrun with the tp70b object
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 200 --warmup 5 > gpurun_out/r2_31_bench_n4.json 2> gpurun_out/r2_31_bench_n4.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_31_bench_n4.json').read().splitlines()[-1]);print(d['metric'],d['value'],d['roofline']['frac']);t=d['tp70b'];print(t['value'],t['comm'],t['roofline'])"; tail -2 gpurun_out/r2_31_bench_n4.err
