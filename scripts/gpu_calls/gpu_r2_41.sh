# round 2, call 41: exchange watchdog -- the single-GPU loop-back tests of the fused GEMV + all-reduce
mkdir -p gpurun_out
timeout -s KILL 70 python -m pytest tests/test_tp_loopback_gpu.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_loopback_tests.txt; cat gpurun_out/r2_loopback_tests.txt
