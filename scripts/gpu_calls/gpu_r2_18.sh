# round 2, call 18: three co-resident CTAs per SM (8 consumer warps, 3-slot rings) against two
mkdir -p gpurun_out
B200AWQ_DECODE_WARPS=83 timeout -s KILL 900 python -m pytest tests/test_gpu_parity_r2.py -q -m gpu -x -k "decode" > gpurun_out/r2_18_tests.log 2>&1; tail -2 gpurun_out/r2_18_tests.log
for shp in "4096 4096" "14336 4096"; do
  timeout -s KILL 200 python scripts/ring_probe.py $shp 1 48 > gpurun_out/r2_18_probe_$(echo $shp | tr ' ' 'x').txt 2>&1
done
grep -h -A2 "^== decode" gpurun_out/r2_18_probe_*.txt | grep -v "^--\|stamps" | cut -c1-170
B200AWQ_DECODE_WARPS=83 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _r18_w83 2>&1 | cut -c1-140
timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _r18_w16 2>&1 | cut -c1-140
