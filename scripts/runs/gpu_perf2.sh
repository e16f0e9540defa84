mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/t_all.log
tail -4 gpurun_out/t_all.log
MS=1,2,4,8,16
timeout -s KILL 300 python scripts/microbench.py --m $MS --tag _v2 > gpurun_out/micro2.log 2>&1
B200AWQ_STREAM_MODE=0 timeout -s KILL 300 python scripts/microbench.py --m 1,4 --tag _v2_mode0 > gpurun_out/micro2_mode0.log 2>&1
B200AWQ_STREAM_RPB=1 timeout -s KILL 300 python scripts/microbench.py --m 1,4 --tag _v2_rpb1 > gpurun_out/micro2_rpb1.log 2>&1
B200AWQ_STREAM_RPB=8 timeout -s KILL 300 python scripts/microbench.py --m 1,4 --tag _v2_rpb8 > gpurun_out/micro2_rpb8.log 2>&1
B200AWQ_PDL=0 timeout -s KILL 300 python scripts/microbench.py --m 1,4 --tag _v2_nopdl > gpurun_out/micro2_nopdl.log 2>&1
B200AWQ_STREAM_KC=2048 timeout -s KILL 300 python scripts/microbench.py --m 1,4 --tag _v2_kc2048 > gpurun_out/micro2_kc2048.log 2>&1
for f in micro2 micro2_mode0 micro2_rpb1 micro2_rpb8 micro2_nopdl micro2_kc2048; do echo == $f; cat gpurun_out/$f.log | cut -c1-110; done
timeout -s KILL 300 python bench.py --steps 200 --warmup 5 --no-cpu --no-ref-gpu > gpurun_out/bench2.json 2> gpurun_out/bench2.err; cut -c1-900 gpurun_out/bench2.json
B200AWQ_PDL=0 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:w4a16_stream -s 3 -c 2 -o gpurun_out/prof_stream_v2 python scripts/prof_one.py 1 4096 4096 > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_stream.log
