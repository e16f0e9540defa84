#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm or one_hot or reference or prefill or large" 2>&1 | tail -4
timeout 200 python scripts/umma_stage_trace.py 2048 4096 4096 0 2>&1 | grep -E "^ +(2[4-9]|6[0-9]|7[0-5]) |steady|tile|=="
timeout 300 python scripts/microbench.py --m 128,256,512,2048 2>&1 | cut -c1-110
