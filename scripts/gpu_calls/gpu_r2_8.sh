# round 2, call 8: MODE 8 with decoupled finisher + REDUX staging: parity (logged), timeline, microbench, bench
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -q -m gpu -x -k "gemv or one_hot or ring or decode or kat" > gpurun_out/r2_8_tests.log 2>&1; tail -8 gpurun_out/r2_8_tests.log
for shp in "4096 4096" "14336 4096" "4096 14336"; do
  timeout -s KILL 200 python scripts/ring_probe.py $shp 1 48 > gpurun_out/r2_8_probe_$(echo $shp | tr ' ' 'x').txt 2>&1
done
grep -A3 "^== int8 digits (mode 8)\|^== int8 digits, dry" gpurun_out/r2_8_probe_*.txt | cut -c1-150
grep "warp 0 \|producer" gpurun_out/r2_8_probe_14336x4096.txt | head -4
timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _ring8b 2>&1 | cut -c1-120
timeout -s KILL 900 python bench.py > gpurun_out/r2_8_bench.json 2> gpurun_out/r2_8_bench.err; python -c "
import json;d=json.load(open('gpurun_out/r2_8_bench.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'])"; tail -2 gpurun_out/r2_8_bench.err
