# round 2, call 3: ring kernel cycle accounting + streaming limit (trace build), then parity of the non-volatile mma build
mkdir -p gpurun_out
for shp in "4096 4096" "14336 4096" "4096 14336"; do
  timeout -s KILL 200 python scripts/ring_probe.py $shp 1 48 > gpurun_out/r2_3_probe_$(echo $shp | tr ' ' 'x').txt 2>&1
done
cat gpurun_out/r2_3_probe_*.txt | cut -c1-150
timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _ring2 2>&1 | cut -c1-120
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -q -m gpu -x -k "gemv or one_hot or ring or decode or kat" 2>&1 | tail -8
