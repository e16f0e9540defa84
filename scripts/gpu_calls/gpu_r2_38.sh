# round 2, call 38: sanity after the clean rebuild (build() from scratch incl. the staged reference call sites)
mkdir -p gpurun_out
timeout -s KILL 150 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout -s KILL 300 python -m pytest tests/test_reference_callsites_gpu.py tests/test_attention_gpu.py tests/test_layer_ops_gpu.py -x -q -m gpu 2>&1 | tail -2
