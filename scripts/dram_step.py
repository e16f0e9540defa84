#!/usr/bin/env python
"""One decode step (160 GEMVs) and one prefill step (160 GEMMs, M = 2048) of bench.py's model as plain C-ABI launches,
for `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:w4a16 -c 320 --csv` (DRAM traffic per launch;
scripts/dram_traffic.py turns the CSV into profiles/traffic.json)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llm_awq_b200 as P  # noqa: E402
from bench import LLAMA3_8B, build_model, G  # noqa: E402

dev = torch.device("cuda", 0)
lib = P.lib()
model = build_model(torch, LLAMA3_8B, torch.float16, dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
for M in (1, 2048):
    xs = {K: torch.randn(M, K, device=dev).half() for K in (4096, 14336)}
    ys = {N: torch.empty(M, N, dtype=torch.float16, device=dev) for N in {m["N"] for m in model}}
    torch.cuda.synchronize()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for m in model:
        a = (p(xs[m["K"]]), p(m["qw"]), p(m["sc"]), p(m["sz"]), p(ys[m["N"]]), M, m["N"], m["K"], G, 0)
        rc = lib.b200awq_w4a16_gemv(*a, st) if M < 8 else lib.b200awq_w4a16_gemm(*a, None, 0, st)
        assert rc == 0, rc
    torch.cuda.synchronize()
print("done")
