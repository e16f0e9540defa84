mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemv or one_hot" 2>&1 | tail -2
timeout -s KILL 300 python scripts/microbench.py --m 1,2 --tag _v11 2>&1 | cut -c1-100
# memory-safety evidence: memcheck over the three kernel families on small shapes (tokens 1..513), racecheck on the decode path
timeout -s KILL 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "small and f16" > gpurun_out/r1_sanitizer_memcheck_gemm_small.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r1_sanitizer_memcheck_gemm_small.log | tail -3
timeout -s KILL 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemv_abi and (24-256 or 512-1024) and f16" > gpurun_out/r1_sanitizer_memcheck_gemv.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r1_sanitizer_memcheck_gemv.log | tail -3
timeout -s KILL 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemv_abi and 64-384 and f16" > gpurun_out/r1_sanitizer_racecheck_gemv.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r1_sanitizer_racecheck_gemv.log | tail -3
