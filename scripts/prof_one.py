#!/usr/bin/env python
"""Launch one shape a few times over a ring of distinct weights (for ncu captures).
    python scripts/prof_one.py M N K [iters]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llm_awq_b200 as P
from scripts.microbench import make_ring
M, N, K = (int(v) for v in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 8
dev = torch.device("cuda", 0)
lib = P.lib()
ring = make_ring(N, K, torch.float16, dev, 256 << 20)
x = (torch.randn(M, K, device=dev) * 0.25).half()
y = torch.empty(M, N, dtype=torch.float16, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
for i in range(iters):
    qw, sc, sz = ring[i % len(ring)]
    if M < 8:
        rc = lib.b200awq_w4a16_gemv(p(x), p(qw), p(sc), p(sz), p(y), M, N, K, 128, 0, st)
    else:
        rc = lib.b200awq_w4a16_gemm(p(x), p(qw), p(sc), p(sz), p(y), M, N, K, 128, 0, None, 0, st)
    assert rc == 0, rc
torch.cuda.synchronize()
print("done", float(y.float().abs().mean()))
