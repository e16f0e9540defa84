# round 2, call 36: first activation loads issued right behind the dependency wait (addresses computed before it)
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity_r2.py tests/test_layer_ops_gpu.py -q -m gpu -x -k "decode or ring_partition or mlp_front" > gpurun_out/r2_36_tests.log 2>&1; tail -2 gpurun_out/r2_36_tests.log
timeout -s KILL 200 python scripts/microbench.py --m 1 --tag _r36 2>&1 | cut -c1-140
for shp in "4096 4096"; do timeout -s KILL 100 python scripts/ring_probe.py $shp 1 48 > gpurun_out/r2_36_probe_$(echo $shp | tr ' ' 'x').txt 2>&1; grep -A2 "^== decode kernel (16" gpurun_out/r2_36_probe_$(echo $shp | tr ' ' 'x').txt | cut -c1-190; done
