mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "full_size or linearity or wqlinear" 2>&1 | tail -15 > gpurun_out/t_fix.log
timeout -s KILL 600 python bench.py --steps 200 --warmup 5 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
timeout -s KILL 600 python scripts/microbench.py --ref > gpurun_out/micro1.log 2>&1
B200AWQ_PDL=0 timeout -s KILL 300 python scripts/microbench.py --m 1,4,16 --tag _nopdl > gpurun_out/micro1_nopdl.log 2>&1
B200AWQ_STREAM_MODE=1 timeout -s KILL 300 python scripts/microbench.py --m 1,4,16 --tag _mode1 > gpurun_out/micro1_mode1.log 2>&1
B200AWQ_STREAM_MODE=2 timeout -s KILL 300 python scripts/microbench.py --m 1,4,16 --tag _mode2 > gpurun_out/micro1_mode2.log 2>&1
tail -5 gpurun_out/t_fix.log; cat gpurun_out/bench1.json; tail -3 gpurun_out/bench1.err; cat gpurun_out/micro1.log
