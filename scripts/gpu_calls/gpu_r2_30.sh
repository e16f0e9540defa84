# round 2, call 30: DRAM traffic of one decode step and one prefill step (ncu metrics pass)
mkdir -p gpurun_out
timeout -s KILL 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:w4a16 -c 320 --csv --log-file gpurun_out/r2_dram_step.csv python scripts/dram_step.py > gpurun_out/r2_dram_step.log 2>&1; tail -2 gpurun_out/r2_dram_step.log; wc -l gpurun_out/r2_dram_step.csv
