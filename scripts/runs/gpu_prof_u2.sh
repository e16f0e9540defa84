mkdir -p gpurun_out
B200AWQ_PDL=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_umma2 -s 2 -c 1 -f -o gpurun_out/r1_umma2_2048x4096x4096 python scripts/prof_one.py 2048 4096 4096 4 > gpurun_out/ncu_d.log 2>&1; tail -1 gpurun_out/ncu_d.log
B200AWQ_UMMA2=0 B200AWQ_PDL=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_umma -s 2 -c 1 -f -o gpurun_out/r1_umma_2048x4096x4096 python scripts/prof_one.py 2048 4096 4096 4 > gpurun_out/ncu_e.log 2>&1; tail -1 gpurun_out/ncu_e.log
