# round 2, call 40: decode attention after the loads-in-flight change -- parity tests, then the same timing as call 39
mkdir -p gpurun_out
timeout -s KILL 100 python -m pytest tests/test_attention_gpu.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_att_tests.txt; cat gpurun_out/r2_att_tests.txt
timeout -s KILL 80 python scripts/layer_ops_bench.py > gpurun_out/r2_layer_ops_b.json 2> gpurun_out/r2_layer_ops_b.err; cat gpurun_out/r2_layer_ops_b.json; tail -3 gpurun_out/r2_layer_ops_b.err
