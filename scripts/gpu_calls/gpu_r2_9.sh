# round 2, call 9 (re-entry): re-measure the ring kernels -- parity, per-warp timelines, microbench per mode, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_9_smi.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -q -m gpu -x -k "gemv or one_hot or ring or decode or kat" > gpurun_out/r2_9_tests.log 2>&1; tail -5 gpurun_out/r2_9_tests.log
for shp in "4096 4096" "14336 4096" "4096 14336" "6144 4096"; do
  timeout -s KILL 200 python scripts/ring_probe.py $shp 1 48 > gpurun_out/r2_9_probe_$(echo $shp | tr ' ' 'x').txt 2>&1
done
grep -h -A3 "^==" gpurun_out/r2_9_probe_*.txt | cut -c1-150
grep "warp 0 \|producer" gpurun_out/r2_9_probe_14336x4096.txt | head -12
timeout -s KILL 300 python scripts/microbench.py --m 1,2,4 --tag _r9_ring 2>&1 | cut -c1-140
B200AWQ_RING_MODE=2 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _r9_ring2 2>&1 | cut -c1-140
B200AWQ_RING=0 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _r9_stream 2>&1 | cut -c1-140
timeout -s KILL 120 scripts/probes/mma_probe.bin > gpurun_out/r2_mma_probe.txt 2>&1; cat gpurun_out/r2_mma_probe.txt | head -40
timeout -s KILL 900 python bench.py > gpurun_out/r2_9_bench.json 2> gpurun_out/r2_9_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2_9_bench.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d.get('reference_on_b200'),d['prefill']['value'],d['prefill']['roofline'])"; tail -2 gpurun_out/r2_9_bench.err
