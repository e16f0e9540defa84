mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -2
timeout -s KILL 300 python scripts/microbench.py --m 1 --tag _v6 2>&1 | cut -c1-100
timeout -s KILL 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; cat gpurun_out/bench_r1.json; tail -2 gpurun_out/bench_r1.err
