mkdir -p gpurun_out
B200AWQ_PDL=0 timeout -s KILL 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:w4a16_stream -s 4 -c 3 -o gpurun_out/prof_stream_v3 -f python scripts/prof_one.py 1 4096 4096 12 > gpurun_out/ncu_stream3.log 2>&1
tail -2 gpurun_out/ncu_stream3.log
B200AWQ_PDL=0 B200AWQ_STREAM_MODE=0 timeout -s KILL 600 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:w4a16_stream -s 4 -c 2 -o gpurun_out/prof_stream_v3m0 -f python scripts/prof_one.py 1 4096 4096 12 > gpurun_out/ncu_stream3m0.log 2>&1
tail -2 gpurun_out/ncu_stream3m0.log
