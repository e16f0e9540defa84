# round 2, call 13: decode kernel, two groups per consumer batch, digit staging with independent chains: parity, timelines, microbench, bench
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_layer_ops_gpu.py -q -m gpu -x -k "gemv or one_hot or ring or decode or kat or rmsnorm" > gpurun_out/r2_13_tests.log 2>&1; tail -5 gpurun_out/r2_13_tests.log
for shp in "4096 4096" "14336 4096" "4096 14336" "6144 4096"; do
  timeout -s KILL 200 python scripts/ring_probe.py $shp 1 48 > gpurun_out/r2_13_probe_$(echo $shp | tr ' ' 'x').txt 2>&1
done
grep -h -A3 "^== decode" gpurun_out/r2_13_probe_*.txt | cut -c1-150
grep -h "^==\|warp 0 \|warp 15\|producer" gpurun_out/r2_13_probe_14336x4096.txt gpurun_out/r2_13_probe_4096x4096.txt | head -40
timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _r13_decode 2>&1 | cut -c1-140
B200AWQ_DECODE_SPLIT=2 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _r13_decode_s2 2>&1 | cut -c1-140
timeout -s KILL 900 python bench.py --no-ref-gpu > gpurun_out/r2_13_bench.json 2> gpurun_out/r2_13_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2_13_bench.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'])"; tail -2 gpurun_out/r2_13_bench.err
