#!/bin/bash
cd /root/repo
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch_sweep or prefill or reference" 2>&1 | tail -2
