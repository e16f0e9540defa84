# round 2, call 29 (closing): the whole GPU suite, smoke(), the bench, ncu captures of the decode and prefill kernels
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/ -q -m gpu > gpurun_out/r2_29_tests.log 2>&1; tail -4 gpurun_out/r2_29_tests.log
timeout -s KILL 200 python __graft_entry__.py --smoke 2>&1 | tail -6
timeout -s KILL 600 python bench.py > gpurun_out/r2_29_bench.json 2> gpurun_out/r2_29_bench.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_29_bench.json').read().splitlines()[-1]);print(len(open('gpurun_out/r2_29_bench.json').read().splitlines()),'line(s);',d['value'],d['roofline']['frac'],d['e2e']['value'],d['e2e_eager']['value'],d['prefill']['value'],d['prefill']['roofline']['frac'])"; tail -2 gpurun_out/r2_29_bench.err
B200AWQ_PDL=0 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:w4a16_decode -s 4 -c 1 -f -o gpurun_out/r2_final_decode_gemv_14336x4096 python scripts/prof_one.py 1 14336 4096 8 > gpurun_out/ncu_a.log 2>&1; tail -1 gpurun_out/ncu_a.log
B200AWQ_PDL=0 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:w4a16_umma -s 2 -c 1 -f -o gpurun_out/r2_final_umma_gemm_2048x14336x4096 python scripts/prof_one.py 2048 14336 4096 5 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:w4a16 -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 --prefill-steps 3 --no-cpu --no-ref-gpu --no-sweep > gpurun_out/r2_launches_bench.log 2>&1; tail -c 300 gpurun_out/r2_launches_bench.log
