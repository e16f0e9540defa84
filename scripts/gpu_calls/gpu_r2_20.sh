# round 2, call 20: attention / RoPE parity first (new kernels under a short timeout), then the whole suite, then the bench
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_attention_gpu.py -q -m gpu -x > gpurun_out/r2_20_att.log 2>&1; tail -15 gpurun_out/r2_20_att.log
timeout -s KILL 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r2_20_tests.log 2>&1; tail -8 gpurun_out/r2_20_tests.log
timeout -s KILL 900 python bench.py > gpurun_out/r2_20_bench.json 2> gpurun_out/r2_20_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_20_bench.json').read().splitlines()[-1]);print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['e2e_eager']['value'],d['e2e_eager']['host_us_per_call']);print(json.dumps(d['reference_on_b200']['decode'])[:700])"; tail -3 gpurun_out/r2_20_bench.err
