# round 2, call 19: whole GPU suite (fused MLP front, RMSNorm, loop-back all-reduce, reference call sites), then the bench
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/r2_19_tests.log 2>&1; tail -6 gpurun_out/r2_19_tests.log
timeout -s KILL 900 python bench.py > gpurun_out/r2_19_bench.json 2> gpurun_out/r2_19_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2_19_bench.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['e2e_eager'],json.dumps(d.get('reference_on_b200'))[:900]);print(json.dumps(d.get('batch_sweep'))[:1500])"; tail -3 gpurun_out/r2_19_bench.err
