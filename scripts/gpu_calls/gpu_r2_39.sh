# round 2, call 39: timing of the ops around the linears against their HBM bounds
mkdir -p gpurun_out
timeout -s KILL 200 python scripts/layer_ops_bench.py > gpurun_out/r2_layer_ops.json 2> gpurun_out/r2_layer_ops.err; cat gpurun_out/r2_layer_ops.json; tail -3 gpurun_out/r2_layer_ops.err
