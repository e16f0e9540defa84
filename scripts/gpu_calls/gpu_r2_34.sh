# round 2, call 34: latency of the first activation loads after the dependency wait (trace build), and parity after the dispatch change
mkdir -p gpurun_out
for shp in "4096 4096" "14336 4096"; do timeout -s KILL 100 python scripts/ring_probe.py $shp 1 48 > gpurun_out/r2_34_probe_$(echo $shp | tr ' ' 'x').txt 2>&1; grep -A3 "^== decode kernel (16" gpurun_out/r2_34_probe_$(echo $shp | tr ' ' 'x').txt | cut -c1-190; grep "SM clock" gpurun_out/r2_34_probe_$(echo $shp | tr ' ' 'x').txt | head -2; done
timeout -s KILL 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -q -m gpu -x -k "gemm or llama2 or sweep or batch" > gpurun_out/r2_34_tests.log 2>&1; tail -2 gpurun_out/r2_34_tests.log
