# round 2, call 24: loop-back of the exchange in the decode kernel's finisher first (short timeout), then the whole suite
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_tp_loopback_gpu.py -q -m gpu -x > gpurun_out/r2_24_loop.log 2>&1; tail -4 gpurun_out/r2_24_loop.log
timeout -s KILL 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r2_24_tests.log 2>&1; tail -5 gpurun_out/r2_24_tests.log
