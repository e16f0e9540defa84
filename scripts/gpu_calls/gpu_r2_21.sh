# round 2, call 21: attention / RoPE parity after the workspace-layout fix
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_attention_gpu.py -q -m gpu > gpurun_out/r2_21_att.log 2>&1; tail -15 gpurun_out/r2_21_att.log | cut -c1-250
