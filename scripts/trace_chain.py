#!/usr/bin/env python
"""Debug experiment: how do consecutive GEMV launches overlap?  Uses the TRACE build of the library
(python llm_awq_b200/build.py --trace) which stamps %globaltimer at 7 points of the first and last CTA.

    python scripts/trace_chain.py [N K M chain]
Modes: plain launches / PDL launches / CUDA graph (+-PDL).  Prints us per launch and, for launches in the
middle of the chain, the stamp times relative to that launch's first stamp and to the previous launch's end.
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scripts.microbench import make_ring  # noqa: E402

N, K, M, CH = (int(v) for v in (sys.argv[1:5] + ["4096", "4096", "1", "48"][len(sys.argv) - 1:]))
lib = ctypes.CDLL(os.path.join(ROOT, "llm_awq_b200", "lib", "libb200awq_trace.so"))
vp, ci = ctypes.c_void_p, ctypes.c_int
lib.b200awq_w4a16_gemv.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
lib.b200awq_w4a16_gemm.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp, ctypes.c_size_t, vp]
lib.b200awq_set_pdl.argtypes = [ci]
lib.b200awq_debug_read_trace.argtypes = [vp, ci]
dev = torch.device("cuda", 0)
ring = make_ring(N, K, torch.float16, dev, 512 << 20)
x = (torch.randn(M, K, device=dev) * 0.25).half()
y = torch.empty(M, N, dtype=torch.float16, device=dev)
p = lambda t: vp(t.data_ptr())


def chain():
    st = vp(torch.cuda.current_stream().cuda_stream)
    for i in range(CH):
        qw, sc, sz = ring[i % len(ring)]
        if M < 8:
            rc = lib.b200awq_w4a16_gemv(p(x), p(qw), p(sc), p(sz), p(y), M, N, K, 128, 0, st)
        else:
            rc = lib.b200awq_w4a16_gemm(p(x), p(qw), p(sc), p(sz), p(y), M, N, K, 128, 0, None, 0, st)
        assert rc == 0, rc


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps / CH


def dump(tag):
    buf = (ctypes.c_ulonglong * (1024 * 16))()
    assert lib.b200awq_debug_read_trace(buf, 1024 * 16) == 0
    t = torch.tensor(list(buf), dtype=torch.int64).view(1024, 2, 8)
    # pick 6 consecutive launches with the largest seq values present (the last chain executed)
    seqs = sorted(range(1024), key=lambda s: int(t[s, 0, 0]))[-CH:]
    seqs = seqs[CH // 2: CH // 2 + 6]
    print(f"[{tag}] ABSOLUTE stamps (us): start | issued | pre-wait | wait-ret | x-ready | loop-end | end | bar-init"
          "   (ring kernel: start | producer issued all | cluster sync done | wait-ret | x staged | consumers done | finisher done)")
    t0 = int(t[seqs[0], 0, 0])
    for s in seqs:
        for c in (0, 1):
            r = t[s, c]
            print(f"  seq {s:4d} cta {'first' if c == 0 else 'last '}: " + " ".join(f"{(int(r[i]) - t0) / 1e3:7.2f}" for i in range(8)))


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for pdl in (0, 1):
        lib.b200awq_set_pdl(pdl)
        us = timed(chain)
        print(f"plain launches pdl={pdl}: {us:.2f} us/launch")
        dump(f"plain pdl={pdl}")
for pdl in (0, 1):
    lib.b200awq_set_pdl(pdl)
    with torch.cuda.stream(s):
        chain()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        chain()
    us = timed(g.replay)
    print(f"graph pdl={pdl}: {us:.2f} us/launch")
    dump(f"graph pdl={pdl}")
