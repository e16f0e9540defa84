"""Shared helpers for the parity tests (synthetic inputs per SURVEY.md §8d)."""
import ctypes

import numpy as np
import torch

from oracle import w4a16_oracle as O


def gen_layer(N, K, dtype=torch.float16, seed=0, device="cpu"):
    """Random packed weights + scales/zeros exactly as WQLinear stores them."""
    g = torch.Generator().manual_seed(seed)
    rows = O.scale_rows(K)
    ng = K // 128
    qweight = torch.randint(-32768, 32768, (N // 4, K), generator=g, dtype=torch.int32).to(torch.int16)
    scales = torch.zeros(rows, N)
    scales[:ng] = 0.004 + 0.012 * torch.rand(ng, N, generator=g)
    scales = scales.to(dtype)
    zint = torch.randint(0, 16, (rows, N), generator=g)
    szeros = (-(scales.float() * zint)).to(dtype)
    szeros[ng:] = 0
    return qweight.to(device), scales.to(device), szeros.to(device)


def gen_x(M, K, dtype=torch.float16, seed=1, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(M, K, generator=g).to(dtype).to(device)


def np64(t):
    return t.detach().float().cpu().numpy().astype(np.float64)


def dt_name(dtype):
    return "f16" if dtype == torch.float16 else "bf16"


def oracle_forward(x, qweight, scales, szeros, dtype, rows=None):
    return O.wq_linear_forward(np64(x), qweight.cpu().numpy(), np64(scales), np64(szeros),
                               dtype=dt_name(dtype), rows=rows)


def rel_err(y, ref):
    y = np.asarray(y, dtype=np.float64)
    return float(np.linalg.norm(y - ref) / max(np.linalg.norm(ref), 1e-30))


def abi_call(lib, x, qweight, scales, szeros, M, N, K, dtype, kind, stream=None):
    """Call the C ABI directly (ctypes, raw device pointers). Returns (code, y)."""
    y = torch.empty(M, N, dtype=dtype, device=x.device)
    code = 0 if dtype == torch.float16 else 1
    st = ctypes.c_void_p(stream if stream is not None else torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    if kind == "gemv":
        rc = lib.b200awq_w4a16_gemv(p(x), p(qweight), p(scales), p(szeros), p(y), M, N, K, 128, code, st)
    else:
        rc = lib.b200awq_w4a16_gemm(p(x), p(qweight), p(scales), p(szeros), p(y), M, N, K, 128, code, None, 0, st)
    return rc, y
