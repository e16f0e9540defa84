# round 2, call 17: producer starts its first lap before the CTA barrier; lane-parallel barrier init; 8 reduction slots
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -q -m gpu -x -k "gemv or one_hot or ring or decode or kat" > gpurun_out/r2_17_tests.log 2>&1; tail -3 gpurun_out/r2_17_tests.log
B200AWQ_DECODE_WARPS=8 timeout -s KILL 900 python -m pytest tests/test_gpu_parity_r2.py -q -m gpu -x -k "decode" > gpurun_out/r2_17_tests8.log 2>&1; tail -2 gpurun_out/r2_17_tests8.log
for shp in "4096 4096" "14336 4096" "4096 14336" "6144 4096"; do
  timeout -s KILL 200 python scripts/ring_probe.py $shp 1 48 > gpurun_out/r2_17_probe_$(echo $shp | tr ' ' 'x').txt 2>&1
done
grep -h -A3 "^== decode" gpurun_out/r2_17_probe_*.txt | grep -v "^--\|stamps\|cta last" | cut -c1-170
grep -h "^==\|warp 0 \|producer" gpurun_out/r2_17_probe_14336x4096.txt gpurun_out/r2_17_probe_4096x4096.txt | head -30
timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _r17_decode 2>&1 | cut -c1-140
B200AWQ_DECODE_WARPS=8 timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _r17_decode_w8 2>&1 | cut -c1-140
