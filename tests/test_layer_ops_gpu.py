"""GPU parity of the ops adjacent to the W4A16 linears: RMSNorm (`layernorm_forward_cuda`) and the
fused gate+up+SiLU*mul front half of the MLP, through the C ABI and the `awq_inference_engine` plugin, against
oracle/layer_oracle.py and -- when present -- the reference's own kernel rebuilt for sm_100a (oracle/_ref).

Tolerances: both ops round ONCE to the element type from fp32, so against the float64 oracle rounded to the type
an element is either equal or one ulp off (fp32 vs fp64 statistics flip a few roundings); the tests bound the
fraction of one-ulp elements and the normwise error (1e-3, north_star)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

import llm_awq_b200 as P
from oracle import layer_oracle as LO
from oracle import w4a16_oracle as O
from awq_testutil import dt_name, gen_layer, gen_x, np64, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
DTYPES = [torch.float16, torch.bfloat16]


def _ulps(y, want, dtype):
    """|y - want| in units of the element type's spacing at `want`."""
    yi = y.view(torch.int16).to(torch.int32)
    wi = want.view(torch.int16).to(torch.int32)
    return (yi - wi).abs()   # same sign and finite: adjacent bit patterns are adjacent values


def _ref_engine():
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        import ref_awq_engine
    except ImportError:
        return None
    return ref_awq_engine if hasattr(ref_awq_engine, "layernorm_forward_cuda") else None


def _rmsnorm_abi(x, g, eps):
    y = torch.empty_like(x)
    m, n = x.numel() // x.shape[-1], x.shape[-1]
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = P.lib().b200awq_rmsnorm(p(x), p(g), p(y), m, n, eps, 0 if x.dtype == torch.float16 else 1,
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, P.lib().b200awq_strerror(rc)
    return y


@pytest.mark.parametrize("dtype", DTYPES, ids=["f16", "bf16"])
@pytest.mark.parametrize("m,n", [(1, 4096), (7, 4096), (2048, 4096), (3, 8192), (64, 14336), (5, 128), (2, 8),
                                 (9, 5120), (4, 11008), (3, 100), (2, 4100), (1, 32768), (2, 40000)])
def test_rmsnorm_vs_oracle(m, n, dtype):
    g = torch.Generator().manual_seed(m * 131 + n)
    x = (torch.randn(m, n, generator=g) * 1.7).to(dtype).to(DEV)
    gamma = (1.0 + 0.2 * torch.randn(n, generator=g)).to(dtype).to(DEV)
    eps = 1e-5
    y = _rmsnorm_abi(x, gamma, eps)
    torch.cuda.synchronize()
    want64 = LO.rmsnorm(np64(x), np64(gamma), eps, dt_name(dtype))
    want = torch.from_numpy(want64).to(dtype).to(DEV)
    u = _ulps(y, want, dtype)
    assert int(u.max()) <= 1
    assert float((u > 0).float().mean()) < 2e-3       # fp32 vs float64 statistics: a few roundings flip
    # against the UNROUNDED result the error is the output rounding itself: 2^-12 / sqrt(3) normwise for fp16,
    # 2^-9 / sqrt(3) = 1.1e-3 .. 1.8e-3 for bf16 (the 1e-3 of north_star is an fp16 figure)
    tol = 1e-3 if dtype == torch.float16 else 4e-3
    assert rel_err(np64(y), LO.rmsnorm(np64(x), np64(gamma), eps, dt_name(dtype), rounded=False)) < tol


@pytest.mark.parametrize("dtype", DTYPES, ids=["f16", "bf16"])
def test_rmsnorm_plugin_signature_and_reference_kernel(dtype):
    """layernorm_forward_cuda(input [b, s, c], gamma, out, eps) -> None (layernorm.cu:111-131, pybind.cpp:17), as
    tinychat/models/llama.py:33-36 calls it; compared with the reference's kernel on the same tensors."""
    eng = P.engine()
    x = (gen_x(2 * 33, 4096, dtype, seed=5, device=DEV) * 2.0).view(2, 33, 4096)
    gamma = (1.0 + 0.1 * gen_x(1, 4096, dtype, seed=6, device=DEV)).view(4096)
    out = torch.empty_like(x)
    assert eng.layernorm_forward_cuda(x, gamma, out, 1e-6) is None
    torch.cuda.synchronize()
    want = torch.from_numpy(LO.rmsnorm(np64(x), np64(gamma), 1e-6, dt_name(dtype))).to(dtype).to(DEV)
    assert int(_ulps(out, want, dtype).max()) <= 1
    ref = _ref_engine()
    if ref is not None:
        out_ref = torch.empty_like(x)
        ref.layernorm_forward_cuda(x, gamma, out_ref, 1e-6)
        torch.cuda.synchronize()
        u = _ulps(out, out_ref, dtype)
        assert int(u.max()) <= 1 and float((u > 0).float().mean()) < 2e-3
    with pytest.raises(RuntimeError):
        eng.layernorm_forward_cuda(x, gamma.float(), out, 1e-6)      # layernorm.cu:122
    with pytest.raises(RuntimeError):
        eng.layernorm_forward_cuda(x.float(), gamma.float(), out.float(), 1e-6)   # fp32 is not dispatched (:14-25)


def test_rmsnorm_fp16_clamp_and_zero_rows():
    """reduction.cuh:78-82: fp16 results are clamped to +-(65504 - 1000) instead of overflowing to inf; an all-zero
    row gives zeros (rsqrt(eps) * 0)."""
    n = 1024
    x = torch.zeros(3, n, dtype=torch.float16, device=DEV)
    x[0, :] = 1.0
    x[2, 5] = 3.0
    gamma = torch.full((n,), 65000.0, dtype=torch.float16, device=DEV)
    y = _rmsnorm_abi(x, gamma, 1e-6)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    assert float(y[0, 0]) == float(torch.tensor(65504.0 - 1000.0).half())
    assert torch.count_nonzero(y[1]) == 0
    assert float(y[2, 5]) == float(torch.tensor(65504.0 - 1000.0).half()) and torch.count_nonzero(y[2]) == 1


def test_rmsnorm_under_graph_replay_in_a_chain():
    """norm -> GEMV -> norm captured in one CUDA graph with programmatic dependent launch: the norm reads its input
    only after the producer finished, and the GEMV after the norm."""
    dtype, H = torch.float16, 4096
    qw, s, z = gen_layer(H, H, dtype, seed=11, device=DEV)
    gamma = (1.0 + 0.1 * gen_x(1, H, dtype, seed=12, device=DEV)).view(H)
    x0 = gen_x(1, H, dtype, seed=13, device=DEV)
    eng = P.engine()
    a, c = torch.empty_like(x0), torch.empty_like(x0)

    def chain():
        eng.layernorm_forward_cuda(x0.view(1, 1, H), gamma, a.view(1, 1, H), 1e-5)
        b = eng.gemv_forward_cuda_new(a, qw, s, z, 1, H, H, 128)
        eng.layernorm_forward_cuda(b.view(1, 1, H), gamma, c.view(1, 1, H), 1e-5)
        return b
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        chain()
    side.synchronize()
    want_c = c.clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        chain()
    for _ in range(4):
        a.zero_(), c.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(c, want_c)
    n1 = LO.rmsnorm(np64(x0), np64(gamma), 1e-5, "f16")
    assert rel_err(np64(a), n1) < 1e-3
