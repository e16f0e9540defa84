#!/bin/bash
# per-warp direct activation fetch in the decode kernel
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not prefill and not large" 2>&1 | tail -3
timeout 300 python scripts/microbench.py --m 1,2,4 --tag _v11 2>&1 | cut -c1-110
timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-400
