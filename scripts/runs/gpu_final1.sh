mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/ -q -m gpu 2>&1 | tail -3
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout -s KILL 600 python scripts/microbench.py --m 1,2,4,8,16,32,64,128,256,2048 --ref --tag _final > gpurun_out/sweep_final.log 2>&1; tail -3 gpurun_out/sweep_final.log | cut -c1-160
timeout -s KILL 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-1200 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout -s KILL 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_cpu.json 2>&1; cut -c1-400 gpurun_out/bench_reference_cpu.json
# per-launch DRAM traffic of one decode step (our kernels only)
timeout -s KILL 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:w4a16 -s 160 -c 160 --csv --log-file gpurun_out/r1_decode_step_dram.csv python bench.py --steps 1 --warmup 3 --prefill-steps 3 --no-cpu --no-ref-gpu > /dev/null 2>&1; wc -l gpurun_out/r1_decode_step_dram.csv
