#!/usr/bin/env python
"""Per-shape microbenchmark of the two kernels (and the reference's, rebuilt for sm_100a).

Every timed launch sequence walks a RING of distinct weight copies whose footprint exceeds the
L2 (>= 512 MB), so each call streams its weights from HBM.  Ours: C-ABI launches captured in a
CUDA graph (one graph = one pass over the ring).  Reference: plain launches (legacy stream).

    python scripts/microbench.py [--m 1,4,16,64,2048] [--dtype f16] [--ref] [--shapes llama3]
Environment knobs of the library are honoured (B200AWQ_STREAM_MODE, B200AWQ_STREAM_KC, B200AWQ_PDL,
B200AWQ_UMMA_TN, B200AWQ_UMMA_CTAS, B200AWQ_STREAM_MAX_M).
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llm_awq_b200 as P  # noqa: E402
from bench import alg_bytes, alg_flops, read_peaks, scale_rows  # noqa: E402

SHAPES = {"llama3": [(6144, 4096), (4096, 4096), (14336, 4096), (4096, 14336)],
          "tp70b": [(1280, 8192), (8192, 1024), (3584, 8192), (8192, 3584)],
          "sq": [(4096, 4096)]}


def make_ring(N, K, dtype, dev, min_bytes):
    per = N * K // 2
    n = max(2, -(-min_bytes // per))
    ring = []
    for i in range(n):
        qw = torch.randint(-32768, 32768, (N // 4, K), dtype=torch.int32, device=dev).to(torch.int16)
        rows = scale_rows(K)
        sc = (0.004 + 0.012 * torch.rand(rows, N, device=dev)).to(dtype)
        sz = (-(sc.float() * torch.randint(0, 16, (rows, N), device=dev))).to(dtype)
        ring.append((qw, sc, sz))
    return ring


def time_graph(fn, iters=20, warm=3):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(warm):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def time_plain(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", default="1,2,4,7,8,16,64,256,2048")
    ap.add_argument("--dtype", default="f16")
    ap.add_argument("--shapes", default="llama3")
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--ring-mb", type=int, default=512)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    dtype = torch.float16 if args.dtype == "f16" else torch.bfloat16
    code = 0 if dtype == torch.float16 else 1
    lib = P.lib()
    peaks = read_peaks()
    ref = None
    if args.ref:
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        import ref_awq_engine as ref  # noqa: E402
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rows = []
    for (N, K) in SHAPES[args.shapes]:
        ring = make_ring(N, K, dtype, dev, args.ring_mb << 20)
        for M in [int(v) for v in args.m.split(",")]:
            x = (torch.randn(M, K, device=dev) * 0.25).to(dtype)
            y = torch.empty(M, N, dtype=dtype, device=dev)
            use_ring = ring if M <= 256 else ring[:4]

            def ours():
                st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                for (qw, sc, sz) in use_ring:
                    if M < 8:
                        rc = lib.b200awq_w4a16_gemv(p(x), p(qw), p(sc), p(sz), p(y), M, N, K, 128, code, st)
                    else:
                        rc = lib.b200awq_w4a16_gemm(p(x), p(qw), p(sc), p(sz), p(y), M, N, K, 128, code, None, 0, st)
                    assert rc == 0, rc
            us = time_graph(ours) * 1e3 / len(use_ring)
            r = {"N": N, "K": K, "M": M, "us": round(us, 3), "GBs": round(alg_bytes(M, N, K) / us / 1e3, 1),
                 "TFs": round(alg_flops(M, N, K) / us / 1e6, 1),
                 "hbm_frac": round(alg_bytes(M, N, K) / us / 1e3 / peaks["hbm"], 3),
                 "tc_frac": round(alg_flops(M, N, K) / us / 1e6 / peaks["tc_burst"], 3)}
            if ref is not None:
                def theirs():
                    for (qw, sc, sz) in use_ring:
                        if M < 8:
                            ref.gemv_forward_cuda_new(x, qw, sc, sz, M, N, K, 128)
                        else:
                            ref.gemm_forward_cuda_new(x, qw, sc, sz)
                with torch.cuda.stream(torch.cuda.default_stream()):
                    rus = time_plain(theirs) * 1e3 / len(use_ring)
                r["ref_us"] = round(rus, 3)
                r["speedup"] = round(rus / us, 2)
            rows.append(r)
            print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"microbench_{args.shapes}_{args.dtype}{args.tag}.json"), "w") as f:
        json.dump({"env": {k: v for k, v in os.environ.items() if k.startswith("B200AWQ")}, "peaks": peaks, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
