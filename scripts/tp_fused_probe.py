#!/usr/bin/env python
"""2+-GPU probe: us per call of the fused GEMV+exchange kernel vs plain GEMV vs GEMV + NCCL (run under torchrun)."""
import ctypes, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import llm_awq_b200 as P
from llm_awq_b200 import tp
from scripts.microbench import make_ring
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
lib = P.lib()
ex = tp.PeerExchange(1, 8192)
p = lambda t: ctypes.c_void_p(t.data_ptr())
for (N, K) in [(8192, 4096), (8192, 1024), (8192, 3584), (8192, 14336)]:
    ring = make_ring(N, K, torch.float16, dev, 512 << 20)
    x = (torch.randn(1, K, device=dev) * 0.25).half()
    y = torch.empty(1, N, dtype=torch.float16, device=dev)
    res = {}
    for mode in ("plain", "nccl", "fused"):
        def chain():
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            for (qw, sc, sz) in ring:
                a = (p(x), p(qw), p(sc), p(sz), p(y), 1, N, K, 128, 0)
                rc = lib.b200awq_w4a16_gemv_allreduce(*a, ex.ptr, st) if mode == "fused" else lib.b200awq_w4a16_gemv(*a, st)
                assert rc == 0, rc
                if mode == "nccl":
                    dist.all_reduce(y)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            chain()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            chain()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) * 1e3 / 10 / len(ring)
    if rank == 0:
        print(f"N={N} K={K} world={world} dbg={os.environ.get('B200AWQ_FUSED_DBG','0')}: " + "  ".join(f"{k} {v:7.2f} us" for k, v in res.items()), flush=True)
dist.destroy_process_group()
