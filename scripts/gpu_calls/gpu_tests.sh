mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gemv or one_hot" 2>&1 | tail -60 > gpurun_out/t_gemv.log
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not gemv and not one_hot" 2>&1 | tail -100 > gpurun_out/t_gemm.log
tail -25 gpurun_out/t_gemv.log; tail -40 gpurun_out/t_gemm.log
