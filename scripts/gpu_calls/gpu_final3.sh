#!/bin/bash
# round-1 closing run: full GPU suite, smoke, batch sweep (ours vs reference kernels), bench line
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/ -q -m gpu 2>&1 | tail -3
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -s KILL 600 python scripts/microbench.py --m 1,2,3,4,8,16,32,64,128,256,512,2048 --ref --tag _final3 > gpurun_out/sweep_final3.log 2>&1; tail -3 gpurun_out/sweep_final3.log | cut -c1-160
timeout -s KILL 900 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; cut -c1-1500 gpurun_out/bench_final3.json; tail -2 gpurun_out/bench_final3.err
