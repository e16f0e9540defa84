"""GPU parity of the decode-step attention and RoPE entries against oracle/layer_oracle.py and a
torch fp32 reference, through the C ABI and the plugin (`single_query_attention`, `fused_rope_with_pos_forward_func`,
same argument order as awq/kernels/csrc/pybind.cpp:24-28, called as tinychat/modules/fused_attn.py:253-254,308-321).

Tolerance: q, k are rounded to T after the rotation exactly like the reference; logits, softmax and the value sum are
fp32 here, float64 in the oracle: 2e-3 normwise (fp16) covers the output rounding and the fp32 sincos / powf."""
import ctypes
import math

import numpy as np
import pytest
import torch

import llm_awq_b200 as P
from oracle import layer_oracle as LO
from awq_testutil import dt_name, np64, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _caches(B, Hkv, D, L, dtype, fill_to, seed):
    g = torch.Generator().manual_seed(seed)
    kc = torch.zeros(B, Hkv, D // 8, L, 8)
    vc = torch.zeros(B, Hkv, L, D)
    kc[:, :, :, :fill_to] = torch.randn(B, Hkv, D // 8, fill_to, 8, generator=g)
    vc[:, :, :fill_to] = torch.randn(B, Hkv, fill_to, D, generator=g)
    return kc.to(dtype).to(DEV), vc.to(dtype).to(DEV)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("B,H,Hkv,D,L,t", [(1, 32, 8, 128, 512, 0), (1, 32, 8, 128, 512, 37), (1, 32, 8, 128, 4096, 2500),
                                           (2, 8, 8, 128, 256, 255), (1, 8, 1, 64, 300, 299), (3, 16, 4, 128, 128, 500),
                                           (1, 64, 8, 128, 1024, 1023), (2, 4, 2, 256, 64, 10)])
def test_single_query_attention_vs_oracle(B, H, Hkv, D, L, t, dtype):
    """Empty cache, short and long contexts (one CTA / split over CTAs), MHA and grouped-query ratios 1-8, a full cache
    and the circular wrap (t >= L), batch > 1 with the QKV views of the fused projection (batch stride > H * D)."""
    eng = P.engine()
    g = torch.Generator().manual_seed(B * 1000 + t)
    qkv = torch.randn(B, 1, H + 2 * Hkv, D, generator=g).to(dtype).to(DEV)        # the fused QKV output, as tinychat views it
    q, k, v = qkv[:, 0, :H], qkv[:, 0, H:H + Hkv], qkv[:, 0, H + Hkv:]
    kc, vc = _caches(B, Hkv, D, L, dtype, min(t, L), seed=t + 1)
    kc64, vc64 = np64(kc), np64(vc)
    out = eng.single_query_attention(q, k, v, kc, vc, None, None, t, D, 500000.0, 1.0, True)
    torch.cuda.synchronize()
    want = LO.single_query_attention(np64(q), np64(k), np64(v), kc64, vc64, t, D, 500000.0, 1.0, True, dt_name(dtype))
    assert out.shape == (B, H, D)
    tol = 2e-3 if dtype == torch.float16 else 8e-3
    assert rel_err(np64(out), want) < tol
    # the caches were updated in place: the value of this step bit for bit; the rotated key to fp32-angle accuracy (the
    # angle t * scale / base^(2i/d) reaches thousands of radians: fp32 carries it to ~1e-4 rad, float64 in the oracle;
    # the reference itself uses fast-math sin / cos); every other slot untouched
    assert np.array_equal(np64(vc), vc64)
    slot = t % L
    got_kc = np64(kc)
    assert np.abs(got_kc[:, :, :, slot] - kc64[:, :, :, slot]).max() <= 2e-3 * np.abs(kc64[:, :, :, slot]).max()
    mask = np.ones(L, dtype=bool)
    mask[slot] = False
    assert np.array_equal(got_kc[:, :, :, mask], kc64[:, :, :, mask])


def test_single_query_attention_decode_loop_matches_sdpa():
    """40 decode steps from an empty cache (the way tinychat drives it) against torch's scaled_dot_product_attention on
    the SAME rotated keys (read back from the cache layout): exercises cache writes feeding later steps."""
    eng = P.engine()
    dtype, B, H, Hkv, D, L = torch.float16, 1, 8, 2, 128, 64
    kc = torch.zeros(B, Hkv, D // 8, L, 8, dtype=dtype, device=DEV)
    vc = torch.zeros(B, Hkv, L, D, dtype=dtype, device=DEV)
    g = torch.Generator().manual_seed(0)
    for t in range(40):
        qkv = torch.randn(B, 1, H + 2 * Hkv, D, generator=g).to(dtype).to(DEV)
        q, k, v = qkv[:, 0, :H], qkv[:, 0, H:H + Hkv], qkv[:, 0, H + Hkv:]
        out = eng.single_query_attention(q, k, v, kc, vc, None, None, t, D, 10000.0, 1.0, True)
        keys = kc[:, :, :, :t + 1].permute(0, 1, 3, 2, 4).reshape(B, Hkv, t + 1, D).float()
        vals = vc[:, :, :t + 1].float()
        qr = torch.from_numpy(LO.rope_rotate(np64(q), t, D, 10000.0, 1.0, True, "f16")).float().to(DEV)
        ref = torch.nn.functional.scaled_dot_product_attention(qr.view(B, H, 1, D), keys.repeat_interleave(H // Hkv, 1),
                                                               vals.repeat_interleave(H // Hkv, 1))
        assert rel_err(np64(out), np64(ref.view(B, H, D))) < 2e-3, t


def test_single_query_attention_length_per_sample_interleaved_rope_and_errors():
    eng = P.engine()
    dtype, B, H, Hkv, D, L = torch.float16, 2, 8, 4, 128, 96
    g = torch.Generator().manual_seed(4)
    q = torch.randn(B, H, D, generator=g).to(dtype).to(DEV)
    k = torch.randn(B, Hkv, D, generator=g).to(dtype).to(DEV)
    v = torch.randn(B, Hkv, D, generator=g).to(dtype).to(DEV)
    kc, vc = _caches(B, Hkv, D, L, dtype, L, seed=9)
    kc64, vc64 = np64(kc), np64(vc)
    lps = torch.tensor([5, 60], dtype=torch.int32, device=DEV)
    out = eng.single_query_attention(q, k, v, kc, vc, lps, None, 77, 64, 10000.0, 0.5, False)   # partial, interleaved rotary
    torch.cuda.synchronize()
    want = LO.single_query_attention(np64(q), np64(k), np64(v), kc64, vc64, 77, 64, 10000.0, 0.5, False, "f16",
                                     length_per_sample=[5, 60])
    assert rel_err(np64(out), want) < 2e-3 and rel_err(np64(kc), kc64) < 1e-3
    with pytest.raises(RuntimeError):
        eng.single_query_attention(q, k, v, kc, vc, None, torch.ones(H, device=DEV), 3, D, 10000.0, 1.0, True)   # ALiBi
    with pytest.raises(RuntimeError):
        eng.single_query_attention(q.float(), k.float(), v.float(), kc.float(), vc.float(), None, None, 3, D, 1e4, 1.0, True)
    with pytest.raises(RuntimeError):
        eng.single_query_attention(q[:, :, :64], k, v, kc, vc, None, None, 3, D, 10000.0, 1.0, True)             # shape


def test_single_query_attention_graph_replay_and_workspace_reuse():
    """Split contexts leave the workspace counters at zero: repeated launches and CUDA-graph replays agree bit for bit."""
    eng = P.engine()
    dtype, B, H, Hkv, D, L, t = torch.float16, 1, 32, 8, 128, 2048, 2000
    g = torch.Generator().manual_seed(1)
    q = torch.randn(B, H, D, generator=g).to(dtype).to(DEV)
    k = torch.randn(B, Hkv, D, generator=g).to(dtype).to(DEV)
    v = torch.randn(B, Hkv, D, generator=g).to(dtype).to(DEV)
    kc, vc = _caches(B, Hkv, D, L, dtype, t, seed=2)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        first = eng.single_query_attention(q, k, v, kc, vc, None, None, t, D, 10000.0, 1.0, True).clone()
        gr = torch.cuda.CUDAGraph()
        side.synchronize()
        with torch.cuda.graph(gr, stream=side):
            o = eng.single_query_attention(q, k, v, kc, vc, None, None, t, D, 10000.0, 1.0, True)
        for _ in range(3):
            gr.replay()
            side.synchronize()
            assert torch.equal(o, first)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("transpose", [True, False])
def test_fused_rope_with_pos_vs_oracle(dtype, transpose):
    """As tinychat calls it: xq viewed [bsz, seqlen, heads, D] out of the fused QKV tensor (strided), freqs
    [seqlen, 1, 1, D] float32 (tinychat/models/llama.py:50-57), transpose_output_memory = True."""
    eng = P.engine()
    bsz, seqlen, H, Hkv, D = 2, 37, 8, 2, 128
    g = torch.Generator().manual_seed(7)
    qkv = torch.randn(bsz, seqlen, H + 2 * Hkv, D, generator=g).to(dtype).to(DEV)
    xq = qkv[:, :, :H]                                                   # [s = bsz, b = seqlen, h, d], strided
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2).float() / D))
    fr = torch.einsum("i,j->ij", torch.arange(seqlen).float(), inv)
    freqs = torch.cat((fr, fr), dim=-1).reshape(seqlen, 1, 1, D)
    freqs_b = freqs.repeat(1, bsz, 1, 1).contiguous().to(DEV)           # index (b_id * S + s_id) * d2: one row per (seq, batch)
    y = eng.fused_rope_with_pos_forward_func(xq, freqs_b, transpose)
    torch.cuda.synchronize()
    assert y.shape == xq.shape
    want = LO.rope_with_pos(np64(xq), np64(freqs_b), dt_name(dtype))
    u = (y.contiguous().view(torch.int16).to(torch.int32) - torch.from_numpy(want).to(dtype).view(torch.int16).to(torch.int32).to(DEV)).abs()
    assert int(u.max()) <= 1 and float((u > 0).float().mean()) < 2e-3    # fp32 sincos vs float64
    if transpose:
        assert y.transpose(0, 1).is_contiguous()
    # partial rotary dimension: the tail is copied
    y2 = eng.fused_rope_with_pos_forward_func(xq, freqs_b[..., :64].contiguous(), False)
    want2 = LO.rope_with_pos(np64(xq), np64(freqs_b[..., :64].contiguous()), dt_name(dtype))
    assert rel_err(np64(y2), want2) < (1e-3 if dtype == torch.float16 else 4e-3)
    assert torch.equal(y2[..., 64:], xq[..., 64:])
