#!/usr/bin/env python
"""Timing probe of the prefill kernel with parts disabled (TRACE build; results are wrong by design)."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scripts.microbench import make_ring
lib = ctypes.CDLL(os.path.join(ROOT, "llm_awq_b200", "lib", "libb200awq_trace.so"))
vp, ci = ctypes.c_void_p, ctypes.c_int
lib.b200awq_w4a16_gemm.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, ctypes.c_size_t, vp]
lib.b200awq_debug_set_umma.argtypes = [ci]
M, N, K = (int(v) for v in (sys.argv[1:4] + ["2048", "4096", "4096"][len(sys.argv) - 1:]))
dev = torch.device("cuda", 0)
ring = make_ring(N, K, torch.float16, dev, 64 << 20)[:4]
x = (torch.randn(M, K, device=dev) * 0.25).half()
y = torch.empty(M, N, dtype=torch.float16, device=dev)
p = lambda t: vp(t.data_ptr())
names = {0: "full kernel", 1: "no dequant ALU", 2: "no tcgen05.st", 3: "no ALU, no st", 4: "no MMA", 8: "no x TMA", 16: "no w TMA",
         7: "no ALU/st/MMA (pipeline skeleton)", 12: "no MMA, no x TMA", 31: "nothing but barriers", 32: "no epilogue", 63: "barriers only, no epilogue", 39: "no ALU/st/MMA/epilogue", 36: "no MMA, no epilogue"}
for dbg in (0, 32, 36, 39, 63, 31, 4):
    assert lib.b200awq_debug_set_umma(dbg) == 0
    def run():
        st = vp(torch.cuda.current_stream().cuda_stream)
        for (qw, sc, sz) in ring:
            rc = lib.b200awq_w4a16_gemm(p(x), p(qw), p(sc), p(sz), p(y), M, N, K, 128, 0, None, 0, st)
            assert rc == 0, rc
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 5 / len(ring)
    print(f"dbg {dbg:2d} {names[dbg]:36s} {us:8.2f} us   ({2*M*N*K/us/1e6:7.1f} TF/s-equivalent)", flush=True)
