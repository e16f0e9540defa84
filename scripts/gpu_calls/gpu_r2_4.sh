# round 2, call 4: two producer warps; ncu --set full of the ring kernel (stall reasons per SASS line); probe again
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -q -m gpu -x -k "gemv_abi or ring or decode" 2>&1 | tail -3
B200AWQ_PDL=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_ring -s 4 -c 1 -f -o gpurun_out/r2_ring_gemv_14336x4096 python scripts/prof_one.py 1 14336 4096 8 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
for shp in "4096 4096" "14336 4096"; do
  timeout -s KILL 200 python scripts/ring_probe.py $shp 1 48 > gpurun_out/r2_4_probe_$(echo $shp | tr ' ' 'x').txt 2>&1
done
grep -A3 "^==" gpurun_out/r2_4_probe_*.txt | cut -c1-150
grep "producer\|consumer 0" gpurun_out/r2_4_probe_14336x4096.txt | head -12
timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _ring3 2>&1 | cut -c1-120
