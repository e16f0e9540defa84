"""The body of the rotary_embedding_neox kernel (llm_awq_b200/csrc/rope_neox.cuh, shared host/device) walked on the
CPU by tests/host_harness/rope_neox_host.cpp and compared with oracle/layer_oracle.py -- indexing, vector / scalar
variants, GQA head counts, partial rotation, token strides, bf16.  The launch itself is covered by the GPU test
(tests/test_zz_rope_neox_gpu.py).  CPU only; builds the harness with g++ into a temporary directory."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import layer_oracle as LO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")

pytestmark = pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_fp16.h")),
                                reason="needs g++ and the CUDA headers")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("rope_neox") / "rope_neox_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "llm_awq_b200", "csrc"),
                    os.path.join(ROOT, "tests", "host_harness", "rope_neox_host.cpp"), "-o", so], check=True)
    lib = ctypes.CDLL(so)
    vp, ci, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    lib.rope_neox_host.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ll, ll, ci, ci, ci]
    lib.rope_neox_host.restype = ci
    return lib


def run(lib, pos, q, k, cache, head_size, q_stride=None, k_stride=None, force_scalar=False, block=96):
    """q / k: torch tensors [tokens, heads, head_size] (possibly views with a larger token stride); rotated in place."""
    tokens, qh, kh = q.shape[0], q.shape[1], k.shape[1]
    dt = {torch.float16: 0, torch.bfloat16: 1}[q.dtype]
    return lib.rope_neox_host(pos.data_ptr(), q.data_ptr(), k.data_ptr(), cache.data_ptr(), tokens, qh, kh, head_size,
                              cache.shape[1], q_stride or q.stride(0), k_stride or k.stride(0), dt, int(force_scalar), block)


def ulp(dtype):
    return {torch.float16: 2.0 ** -10, torch.bfloat16: 2.0 ** -7}[dtype]


def close(got, want64, dtype):
    """one rounding of fp32 products against one rounding of float64 products: <= 1 ulp of T."""
    g = got.double().numpy()
    return np.all(np.abs(g - want64) <= ulp(dtype) * np.maximum(np.abs(want64), 2.0 ** -14))


CASES = [  # tokens, q_heads, k_heads, head_size, rot_dim
    (5, 4, 4, 16, 16), (3, 8, 2, 64, 64), (7, 32, 8, 128, 128), (2, 3, 1, 24, 16), (4, 2, 2, 12, 6), (1, 1, 0, 8, 8), (6, 5, 5, 40, 40),
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("force_scalar", [False, True])
def test_kernel_body_on_the_host_matches_oracle(harness, case, dtype, force_scalar):
    tokens, qh, kh, hs, rot = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    name = "f16" if dtype == torch.float16 else "bf16"
    maxpos = 97
    cache = torch.from_numpy(LO.neox_cos_sin_cache(rot, maxpos, 10000.0, name)).to(dtype).contiguous()
    pos = torch.randint(0, maxpos, (tokens,), generator=g)
    q = torch.randn(tokens, qh, hs, generator=g).to(dtype)
    k = torch.randn(tokens, max(kh, 1), hs, generator=g).to(dtype)[:, :kh]
    k = k.contiguous() if kh else torch.empty(tokens, 0, hs, dtype=dtype)
    want_q, want_k = LO.rotary_embedding_neox(pos.numpy(), q.double().numpy(), k.double().numpy(), hs, cache.double().numpy(), "f64")
    q0 = q.clone()
    used = run(harness, pos, q, k, cache, hs, force_scalar=force_scalar)
    vectorisable = (rot // 2) % 8 == 0 and hs % 8 == 0
    assert used == (8 if vectorisable and not force_scalar else 1)
    assert close(q, want_q, dtype) and close(k, want_k, dtype)
    assert torch.equal(q[..., rot:], q0[..., rot:])                        # beyond rot_dim: untouched bit for bit
    if tokens and not bool((pos == 0).all()):
        assert not torch.equal(q, q0)


def test_token_strides_and_fused_qkv_views(harness):
    """query and key as views into one [tokens, (q_heads + 2 k_heads) * head_size] buffer (a fused qkv output): the
    token stride is the row of the fused buffer and everything outside the two views stays untouched."""
    g = torch.Generator().manual_seed(9)
    tokens, qh, kh, hs = 6, 8, 2, 64
    row = (qh + 2 * kh) * hs
    buf = torch.randn(tokens, row, generator=g).half()
    before = buf.clone()
    q = buf[:, : qh * hs].view(tokens, qh, hs)
    k = buf[:, qh * hs: (qh + kh) * hs].view(tokens, kh, hs)
    cache = torch.from_numpy(LO.neox_cos_sin_cache(hs, 50)).half()
    pos = torch.randint(0, 50, (tokens,), generator=g)
    want_q, want_k = LO.rotary_embedding_neox(pos.numpy(), q.double().numpy(), k.double().numpy(), hs, cache.double().numpy(), "f64")
    assert run(harness, pos, q, k, cache, hs, q_stride=row, k_stride=row) == 8
    assert close(q, want_q, torch.float16) and close(k, want_k, torch.float16)
    assert torch.equal(buf[:, (qh + kh) * hs:], before[:, (qh + kh) * hs:])  # the value heads


def test_block_size_does_not_matter(harness):
    g = torch.Generator().manual_seed(3)
    cache = torch.from_numpy(LO.neox_cos_sin_cache(32, 20)).half()
    pos = torch.randint(0, 20, (4,), generator=g)
    q0, k0 = torch.randn(4, 6, 32, generator=g).half(), torch.randn(4, 2, 32, generator=g).half()
    outs = []
    for block in (32, 96, 512):
        q, k = q0.clone(), k0.clone()
        run(harness, pos, q, k, cache, 32, block=block)
        outs.append((q, k))
    for q, k in outs[1:]:
        assert torch.equal(q, outs[0][0]) and torch.equal(k, outs[0][1])
