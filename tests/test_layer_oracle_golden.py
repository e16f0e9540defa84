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
