"""oracle/layer_oracle.py against what the reference's own Python gives on CPU (tests/golden/reference_layer_ops.npz,
produced by tests/golden/make_golden.py::layer_ops_golden from tinychat/models/llama.py and the tensor expressions of
tinychat/modules/fused_attn.py:256-305).  The reference's kernels for these ops are CUDA-only; these fixtures pin the
restatement's conventions -- RoPE pairing and angles, the frequency table's layout, the KV-cache layout, the softmax
attention, the RMSNorm formula -- to the definitions the reference keeps beside its kernels.  CPU only."""
import os

import numpy as np
import pytest

from oracle import layer_oracle as LO

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "reference_layer_ops.npz"))
N_NORM = 3
N_ATT = 3
# the fixtures come from fp32 complex arithmetic (apply_rotary_emb) or fp64 (RMSNorm, attention)
RTOL32 = 2e-6


def cfg(i):
    S, H, Hkv, D, theta, scale = GOLD[f"a{i}_cfg"]
    return int(S), int(H), int(Hkv), int(D), float(theta), float(scale)


@pytest.mark.parametrize("i", range(N_NORM))
def test_rmsnorm_matches_reference_definition(i):
    x, w, eps, y = (GOLD[f"n{i}_{k}"] for k in ("x", "w", "eps", "y"))
    got = LO.rmsnorm(x, w, float(eps), rounded=False)
    np.testing.assert_allclose(got, y, rtol=1e-12, atol=0)


@pytest.mark.parametrize("i", range(N_ATT))
def test_rope_neox_matches_apply_rotary_emb(i):
    S, H, Hkv, D, theta, scale = cfg(i)
    for name, heads in (("q", H), ("k", Hkv)):
        x, r = GOLD[f"a{i}_x{name}"][0], GOLD[f"a{i}_r{name}"][0]          # [S, heads, D]
        for t in range(S):
            got = LO.rope_rotate(x[t], t, D, theta, scale, True, "f64")
            np.testing.assert_allclose(got, r[t], rtol=RTOL32, atol=RTOL32 * np.abs(r[t]).max())


@pytest.mark.parametrize("i", range(N_ATT))
def test_interleaved_rope_is_the_same_rotation_on_adjacent_pairs(i):
    """neox = False pairs (2j, 2j + 1) with the angle of pair j: permuting a neox vector into interleaved order,
    rotating, and permuting back is the neox rotation (the reference template treats both the same way,
    decoder_masked_multihead_attention_template.hpp:1080-1135)."""
    S, H, Hkv, D, theta, scale = cfg(i)
    x = GOLD[f"a{i}_xq"][0]
    perm = np.empty(D, dtype=np.int64)
    perm[0::2], perm[1::2] = np.arange(D // 2), np.arange(D // 2) + D // 2
    for t in (0, S // 2, S - 1):
        a = LO.rope_rotate(x[t], t, D, theta, scale, True, "f64")
        b = LO.rope_rotate(x[t][..., perm], t, D, theta, scale, False, "f64")
        np.testing.assert_allclose(b, a[..., perm], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("i", range(N_ATT))
def test_rope_with_pos_matches_apply_rotary_emb_on_the_reference_table(i):
    """fused_rope_with_pos consumes precompute_freqs' table (always unscaled, llama.py:50-57); with that table it is
    apply_rotary_emb at scale 1."""
    S, H, Hkv, D, theta, scale = cfg(i)
    table = GOLD[f"a{i}_table"]                                            # [S, D] = cat(freqs, freqs)
    assert np.array_equal(table[:, : D // 2], table[:, D // 2:])
    for name in ("q", "k"):
        x = GOLD[f"a{i}_x{name}"].transpose(1, 0, 2, 3)                    # [S, B = 1, heads, D]
        want = GOLD[f"a{i}_r{name}1"].transpose(1, 0, 2, 3)
        got = LO.rope_with_pos(x, table, "f64")
        np.testing.assert_allclose(got, want, rtol=RTOL32, atol=RTOL32 * np.abs(want).max())


@pytest.mark.parametrize("i", range(N_ATT))
@pytest.mark.parametrize("ring", [False, True])
def test_decode_step_is_the_last_row_of_the_prefill(i, ring):
    """The cache as the prefill branch stores it (fused_attn.py:262-270) for positions 0 .. S-2, then ONE decode step
    at timestep S-1 with the raw q / k / v of the last position: the oracle must write the cache slot the prefill
    would have written and return the last row of the prefill's causal attention (fused_attn.py:290-304).
    ``ring``: the same through a cache of S - 2 slots that has wrapped (positions stored at p % L)."""
    S, H, Hkv, D, theta, scale = cfg(i)
    ck, cv = GOLD[f"a{i}_cache_k"], GOLD[f"a{i}_cache_v"]                  # [1, Hkv, D/8, S, 8], [1, Hkv, S, D]
    xq, xk, xv = (GOLD[f"a{i}_x{n}"][:, S - 1] for n in "qkv")             # [1, heads, D]
    want = GOLD[f"a{i}_out"][:, S - 1]                                     # [1, H, D]
    if not ring:
        L = S
        k_cache, v_cache = ck.copy(), cv.copy()
        k_cache[:, :, :, S - 1, :] = 0
        v_cache[:, :, S - 1, :] = 0
        t = S - 1
    else:
        # a cache of L = S - 2 slots has wrapped by the last position: positions 0 and 1 have been overwritten by
        # S - 2 and (in this step) S - 1, and exactly the last L positions must be attended to
        L = S - 2
        k_cache = np.zeros((1, Hkv, D // 8, L, 8))
        v_cache = np.zeros((1, Hkv, L, D))
        for p in range(S - 1):                                              # later positions overwrite earlier ones
            k_cache[:, :, :, p % L, :] = ck[:, :, :, p, :]
            v_cache[:, :, p % L, :] = cv[:, :, p, :]
        t = S - 1
    got = LO.single_query_attention(xq, xk, xv, k_cache, v_cache, t, D, theta, scale, True, "f64")
    slot = t % L
    np.testing.assert_allclose(k_cache[:, :, :, slot, :], ck[:, :, :, S - 1, :], rtol=RTOL32,
                               atol=RTOL32 * np.abs(ck).max())
    np.testing.assert_array_equal(v_cache[:, :, slot, :], cv[:, :, S - 1, :])
    if not ring:
        np.testing.assert_allclose(got, want, rtol=5e-6, atol=5e-6 * np.abs(want).max())
        return
    # ring: positions S - L .. S - 1 only; rebuild that answer from the fixture's rotated tensors
    rq, rk = GOLD[f"a{i}_rq"][0, S - 1], GOLD[f"a{i}_rk"][0]               # [H, D], [S, Hkv, D]
    xv_all = GOLD[f"a{i}_xv"][0]
    G = H // Hkv
    ref = np.zeros((H, D))
    for h in range(H):
        K = rk[S - L:, h // G]
        V = xv_all[S - L:, h // G]
        s = K @ rq[h] / np.sqrt(D)
        p = np.exp(s - s.max())
        ref[h] = (p / p.sum()) @ V
    np.testing.assert_allclose(got[0], ref, rtol=5e-6, atol=5e-6 * np.abs(ref).max())


N_CACHE = 3


@pytest.mark.parametrize("i", range(N_CACHE))
def test_neox_table_and_rotation_match_the_reference_python(i):
    """The cos|sin table QuantLlamaRotaryEmbedding builds (fused_attn.py:43-59, fp32 -> fp16) against the oracle's
    (float64 -> fp16: at most the last bit differs), and the oracle's rotary_embedding_neox on the REFERENCE's table
    against apply_rotary_emb at the same positions (differs by the table's fp16 rounding only)."""
    dim, maxpos, base, heads = (int(v) for v in GOLD[f"c{i}_cfg"])
    cache = GOLD[f"c{i}_cache"].astype(np.float64)
    assert cache.shape == (maxpos, dim)
    mine = LO.neox_cos_sin_cache(dim, maxpos, float(base), "f16")
    # fp32 sin / cos of angles up to maxpos radians carry ~maxpos * 2^-24 of argument error before the fp16 rounding
    assert np.abs(mine - cache).max() <= 2.0 ** -11 + maxpos * 2.0 ** -23
    assert (mine != cache).mean() < 0.02
    pos = GOLD[f"c{i}_pos"]
    xq, xk = GOLD[f"c{i}_xq"][0], GOLD[f"c{i}_xk"][0]                      # [tokens, heads, dim]
    rq, rk = LO.rotary_embedding_neox(pos, xq, xk, dim, cache, "f64")
    for got, want in ((rq, GOLD[f"c{i}_rq"][0]), (rk, GOLD[f"c{i}_rk"][0])):
        assert np.abs(got - want).max() <= 2 * 2.0 ** -11 * np.abs(want).max() + 1e-6


def test_neox_rotation_leaves_the_tail_of_a_head_alone():
    rng = np.random.default_rng(5)
    cache = LO.neox_cos_sin_cache(8, 16)                                     # rot_dim 8 < head_size 12
    q, k = rng.standard_normal((3, 2, 12)), rng.standard_normal((3, 1, 12))
    rq, rk = LO.rotary_embedding_neox(np.array([1, 15, 0]), q, k, 12, cache, "f64")
    np.testing.assert_array_equal(rq[..., 8:], q[..., 8:])
    np.testing.assert_array_equal(rk[..., 8:], k[..., 8:])
    np.testing.assert_array_equal(rq[2], q[2])                               # position 0: identity
    assert not np.allclose(rq[0, :, :8], q[0, :, :8])


MLP = np.load(os.path.join(HERE, "golden", "reference_mlp_callsite.npz"))


@pytest.mark.parametrize("tokens", [1, 3, 7, 8, 9, 33])
@pytest.mark.parametrize("i", [0, 1])
def test_mlp_front_matches_the_reference_call_site_executed_on_cpu(i, tokens):
    """tests/golden/reference_mlp_callsite.npz = the reference's unmodified QuantLlamaMLP.our_llama_mlp run on CPU over
    oracle-backed engine calls (make_golden.py::mlp_callsite_golden).  mlp_front must reproduce its result: same
    branch rule (< 8 tokens), same zeros in the GEMM branch (scaled_zeros - 8 * scales in fp16, fused_mlp.py:69,76),
    roundings in the same places.  torch's fp16 silu goes through fp32, the oracle through float64: a rare last-bit
    difference of silu is allowed (and then propagates through the product as at most one more ulp)."""
    from oracle import w4a16_oracle as O
    x = MLP[f"m{i}_t{tokens}_x"]
    gate = [MLP[f"m{i}_gate_{k}"] for k in ("qweight", "scales", "szeros")]
    up = [MLP[f"m{i}_up_{k}"] for k in ("qweight", "scales", "szeros")]
    if tokens >= 8:
        for trip, name in ((gate, "gate"), (up, "up")):
            passed = O.rn_f16(trip[2].astype(np.float64) - O.rn_f16(8.0 * trip[1].astype(np.float64)))
            np.testing.assert_array_equal(passed, MLP[f"m{i}_t{tokens}_{name}_zeros_passed"].astype(np.float64))
            trip[2] = passed
    got = LO.mlp_front(x, gate, up, "f16")
    want = MLP[f"m{i}_t{tokens}_c"].astype(np.float64)
    assert got.shape == want.shape
    diff = np.abs(got - want)
    ulp = 2.0 ** -10 * np.maximum(np.abs(want), 2.0 ** -14)
    assert np.all(diff <= 2 * ulp)
    assert (diff > 0).mean() <= 0.01
