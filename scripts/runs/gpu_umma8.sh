#!/bin/bash
# merged A-ring/smem-ring protocol: parity first, then probe + microbench
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm or one_hot or reference or prefill or large" 2>&1 | tail -4
timeout 200 python scripts/umma_probe.py 2048 4096 4096 2>&1 | tail -8
timeout 300 python scripts/microbench.py --m 32,64,128,256,512,2048 2>&1 | tail -40
