#!/bin/bash
# same-box A/B/H: H = committed prefill kernel, A = + next-stage wait taken between the MMAs, B = A + next tile's scales requested before the epilogue
cd /root/repo
L=llm-awq_b200/lib
for v in H A B H A B; do
  cp $L/variant_$v/libb200awq.so $L/libb200awq.so
  echo "== variant $v"
  timeout 300 python scripts/microbench.py --m 512,2048 --tag _u11$v 2>&1 | cut -c1-72
done
for v in A B; do
  cp $L/variant_$v/libb200awq.so $L/libb200awq.so
  echo "== tests variant $v"
  timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm or one_hot or reference or prefill or large or batch_sweep" 2>&1 | tail -2
done
echo "== stage trace (A)"
timeout 200 python scripts/umma_stage_trace.py 2048 4096 4096 0 2>&1 | grep -E "^ +(2[4-9]|3[0-5]|6[0-9]) |steady|tile|=="
