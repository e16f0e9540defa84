mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --workload tp70b --gpus 1 --steps 50 --warmup 5 > gpurun_out/bench_tp70b_n1.json 2> gpurun_out/err_tp1.log; cut -c1-400 gpurun_out/bench_tp70b_n1.json; tail -2 gpurun_out/err_tp1.log
timeout -s KILL 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['traffic'],d['prefill']['value'],d['prefill']['e2e']['value'],d['prefill']['roofline']['frac'],d['prefill']['roofline']['frac_of_burst'],d['clocks'])"
