#!/usr/bin/env python
"""Where does the persistent ring decode kernel spend its time?  (trace build: python llm_awq_b200/build.py --trace)

    python scripts/ring_probe.py [N K [M [chain]]]
For a chain of launches over distinct weights replayed from a CUDA graph it prints us per launch, the %globaltimer
stamps of two launches in the middle of the chain, and the per-warp cycle accounting of CTA 0 (cycles waiting for a
ring slot / in the group steps / in the cross-warp hand-off), for the normal kernel and with the consumers' math
switched off (the streaming limit of the ring itself), for several ring depths.
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
lib.b200awq_debug_read_trace.argtypes = [vp, ci]
lib.b200awq_debug_ring_stats.argtypes = [vp]
lib.b200awq_debug_ring_flags.argtypes = [ci]
dev = torch.device("cuda", 0)
ring = make_ring(N, K, torch.float16, dev, 512 << 20)
x = (torch.randn(M, K, device=dev) * 0.25).half()
y = torch.empty(M, N, dtype=torch.float16, device=dev)
p = lambda t: vp(t.data_ptr())


def chain():
    st = vp(torch.cuda.current_stream().cuda_stream)
    for i in range(CH):
        qw, sc, sz = ring[i % len(ring)]
        rc = lib.b200awq_w4a16_gemv(p(x), p(qw), p(sc), p(sz), p(y), M, N, K, 128, 0, st)
        assert rc == 0, rc


def run(tag, flags, env):
    for k in ("B200AWQ_RING_SLOTS", "B200AWQ_RING_SPLIT", "B200AWQ_RING_MODE", "B200AWQ_DECODE", "B200AWQ_DECODE_SPLIT",
              "B200AWQ_DECODE_SLOTS", "B200AWQ_DECODE_WARPS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    lib.b200awq_reload_config()
    lib.b200awq_debug_ring_flags(flags)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        chain()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        chain()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 10 / CH
    gbs = (N * K / 2 + 4 * (K // 128) * N) / us / 1e3
    print(f"== {tag}: {us:.2f} us/launch  ({gbs:.0f} GB/s)")
    buf = (ctypes.c_ulonglong * (1024 * 16))()
    assert lib.b200awq_debug_read_trace(buf, 1024 * 16) == 0
    t = torch.tensor(list(buf), dtype=torch.int64).view(1024, 2, 8)
    seqs = sorted(range(1024), key=lambda q: int(t[q, 0, 0]))[-CH:]
    seqs = seqs[CH // 2: CH // 2 + 2]
    t0 = int(t[seqs[0], 0, 0])
    print("   stamps (us): start | producer done | cluster sync | wait-ret | x staged | consumers done | finisher done")
    for q in seqs:
        for c in (0, 1):
            r = t[q, c]
            print(f"   seq {q:4d} cta {'first' if c == 0 else 'last '}: " + " ".join(f"{(int(r[i]) - t0) / 1e3:7.2f}" for i in range(7)) + f"   x loaded {(int(r[7]) - t0) / 1e3:7.2f}")
    st = (ctypes.c_longlong * 80)()
    assert lib.b200awq_debug_ring_stats(st) == 0
    st = list(st)
    if st[77] > 0:
        print(f"   SM clock over the producer's loop: {st[76]} cycles in {st[77]} ns = {st[76] / st[77] * 1e3:.0f} MHz;"
              f"  warp 0: dependency wait -> first activation loads issued {st[79]} cycles, issue -> data {st[78]} cycles")
    for w in range(19):
        a, b, c, n = st[4 * w:4 * w + 4]
        role = "producer" if w == 18 else f"warp {w}"
        if n and (w < 2 or w >= 8):
            print(f"   {role:11s}: wait {a:7d}  work {b:7d}  hand-off {c:7d} cycles over {n} slots  (work/slot {b // n})")


if M == 1:
    run("decode kernel (16 consumer warps)", 0, {})
    run("decode kernel, 8 consumer warps", 0, {"B200AWQ_DECODE_WARPS": "8"})
    run("decode kernel, dry (no math)", 1, {})
    run("decode kernel, k split 2", 0, {"B200AWQ_DECODE_SPLIT": "2"})
    run("decode kernel, no k split", 0, {"B200AWQ_DECODE_SPLIT": "1"})
run("ring kernel, fp16 MACs (w4a16_ring.cu)", 0, {"B200AWQ_DECODE": "0"})
