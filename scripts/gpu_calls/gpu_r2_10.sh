# round 2, call 10: HBM streaming probe per load path; decode parity; microbench per mode; bench; rmsnorm tests
mkdir -p gpurun_out
timeout -s KILL 120 scripts/probes/hbm_probe.bin > gpurun_out/r2_hbm_probe.txt 2>&1; cat gpurun_out/r2_hbm_probe.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_layer_ops_gpu.py -q -m gpu -k "gemv or one_hot or ring or decode or kat or rmsnorm" > gpurun_out/r2_10_tests.log 2>&1; tail -15 gpurun_out/r2_10_tests.log
timeout -s KILL 300 python scripts/microbench.py --m 1,2,4 --tag _r10_ring 2>&1 | cut -c1-140
B200AWQ_RING_MODE=2 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _r10_ring2 2>&1 | cut -c1-140
B200AWQ_RING=0 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _r10_stream 2>&1 | cut -c1-140
timeout -s KILL 900 python bench.py > gpurun_out/r2_10_bench.json 2> gpurun_out/r2_10_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2_10_bench.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d.get('reference_on_b200'),d['prefill']['value'],d['prefill']['roofline'])"; tail -2 gpurun_out/r2_10_bench.err
