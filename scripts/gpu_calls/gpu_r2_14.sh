# round 2, call 14: ncu --set full of the decode kernel (stall reasons per SASS line), isolated launches
mkdir -p gpurun_out
B200AWQ_PDL=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_decode -s 4 -c 1 -f -o gpurun_out/r2_decode_gemv_14336x4096 python scripts/prof_one.py 1 14336 4096 8 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
B200AWQ_PDL=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_decode -s 4 -c 1 -f -o gpurun_out/r2_decode_gemv_4096x4096 python scripts/prof_one.py 1 4096 4096 8 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
