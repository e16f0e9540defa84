# round 2, call 37 (final): the whole GPU suite, smoke(), the bench
mkdir -p gpurun_out
timeout -s KILL 700 python -m pytest tests/ -q -m gpu > gpurun_out/r2_37_tests.log 2>&1; tail -3 gpurun_out/r2_37_tests.log
timeout -s KILL 150 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout -s KILL 500 python bench.py > gpurun_out/r2_37_bench.json 2> gpurun_out/r2_37_bench.err; python -c "
import json;L=open('gpurun_out/r2_37_bench.json').read().splitlines();d=json.loads(L[-1]);print(len(L),'line(s);',d['value'],d['roofline']['frac'],d['roofline']['traffic'],d['e2e']['value'],d['e2e_eager']['value'],d['prefill']['value'],d['prefill']['roofline']['frac'],d['prefill']['roofline']['traffic']);print({k:v['frac'] for k,v in d['batch_sweep'].items()})"
