#!/usr/bin/env python
"""Per-stage clock64 timeline of CTA 0 of the prefill kernel (TRACE build): who waits for whom in the
TMA -> dequant -> tcgen05.st -> MMA -> commit ring.  Usage: umma_stage_trace.py [M N K] [dbg]"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scripts.microbench import make_ring
lib = ctypes.CDLL(os.path.join(ROOT, "llm_awq_b200", "lib", "libb200awq_trace.so"))
vp, ci = ctypes.c_void_p, ctypes.c_int
lib.b200awq_w4a16_gemm.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, ctypes.c_size_t, vp]
lib.b200awq_debug_set_umma.argtypes = [ci]
lib.b200awq_debug_read_umma_trace.argtypes = [vp]
args = [int(v) for v in sys.argv[1:]]
M, N, K = (args[:3] + [2048, 4096, 4096][len(args[:3]):])
dbgs = args[3:] or [0]
dev = torch.device("cuda", 0)
(qw, sc, sz), = make_ring(N, K, torch.float16, dev, 64 << 20)[:1]
x = (torch.randn(M, K, device=dev) * 0.25).half()
y = torch.empty(M, N, dtype=torch.float16, device=dev)
p = lambda t: vp(t.data_ptr())
EV = ["P:empty", "P:tma", "D:full", "D:afull", "M:afull", "M:commit"]
for dbg in dbgs:
    assert lib.b200awq_debug_set_umma(dbg) == 0
    for _ in range(3):
        rc = lib.b200awq_w4a16_gemm(p(x), p(qw), p(sc), p(sz), p(y), M, N, K, 128, 0, None, 0, vp(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
    torch.cuda.synchronize()
    buf = np.zeros((8, 256), dtype=np.int64)
    assert lib.b200awq_debug_read_umma_trace(buf.ctypes.data) == 0
    nst = min(256, (K // 64) * 2)
    t0 = buf[0, 0]
    print(f"== dbg {dbg}: M {M} N {N} K {K}; cycles relative to the producer's first stage")
    print("stage " + " ".join(f"{e:>9s}" for e in EV) + "   d(M:commit)")
    prev = None
    for i in list(range(0, 12)) + list(range(24, 44)) + list(range(K // 64 - 4, min(nst, K // 64 + 12))):
        row = buf[:6, i] - t0
        d = "" if prev is None else f"{buf[5, i] - prev:6d}"
        prev = buf[5, i]
        print(f"{i:5d} " + " ".join(f"{v:9d}" for v in row) + "   " + d)
    st = buf[5, 8:K // 64]
    print("steady-state commit-to-commit cycles: mean %.0f  median %.0f" % (np.diff(st).mean(), np.median(np.diff(st))))
    for t in range(2):
        print(f"tile {t}: dfull seen at {buf[6, t] - t0}, epilogue done at {buf[7, t] - t0}  (epilogue {buf[7, t] - buf[6, t]} cycles)")
