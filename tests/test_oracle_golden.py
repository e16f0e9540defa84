"""The oracle (oracle/w4a16_oracle.py) against vectors produced by the reference's own
Python (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import w4a16_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_packer.npz"))


@pytest.mark.parametrize("idx", range(6))
def test_pack_matches_reference(idx):
    q = G[f"pack{idx}_q"]
    ref = G[f"pack{idx}_packed"]
    mine = O.pack_intweight(q)
    assert mine.dtype == np.int16 and mine.shape == ref.shape
    assert np.array_equal(mine, ref)
    assert np.array_equal(O.unpack_intweight(ref), q)


def test_zeros_width_table():
    for i, a, b in zip(G["zw_in"], G["zw_g128"], G["zw_g64"]):
        assert O.calculate_zeros_width(int(i), 128) == a
        assert O.calculate_zeros_width(int(i), 64) == b


def test_init_shapes():
    for K, N, qr, qc, sr, sc, zr, zc in G["init_shapes"]:
        assert (qr, qc) == (N // 4, K)
        assert (sr, sc) == (O.scale_rows(int(K)), N)
        assert (zr, zc) == (O.scale_rows(int(K)), N)


@pytest.mark.parametrize("tag", ["f32", "f16", "bf16"])
def test_quantiser_and_from_linear(tag):
    w0 = G[f"fl_{tag}_w0"]
    if tag == "f32":  # the quantiser formulas are only bit-reproducible in fp32 arithmetic
        s, z = O.pseudo_quantize_scales_zeros(w0, 128)
        assert np.array_equal(s, G[f"fl_{tag}_scales"])
        assert np.array_equal(z, G[f"fl_{tag}_zeros"])
    qw, sbuf, zbuf = O.from_linear(G[f"fl_{tag}_wdq"], G[f"fl_{tag}_scales"], G[f"fl_{tag}_zeros"], 128, tag)
    assert np.array_equal(qw, G[f"fl_{tag}_qweight"])
    assert np.array_equal(sbuf, G[f"fl_{tag}_sbuf"])
    assert np.array_equal(zbuf, G[f"fl_{tag}_zbuf"])


@pytest.mark.parametrize("tag", ["f16", "bf16"])
def test_dequant_reproduces_reference_fake_quant(tag):
    """oracle dequant of the packed tensors ~= the reference's simulated-quant weight
    (quantizer.py:93-95), up to the two roundings of the packed representation."""
    wdq = G[f"fl_{tag}_wdq"].astype(np.float64)
    w = O.dequant_weight(G[f"fl_{tag}_qweight"], G[f"fl_{tag}_sbuf"], G[f"fl_{tag}_zbuf"], 128, tag)
    eps = 2.0 ** -10 if tag == "f16" else 2.0 ** -7
    tol = eps * (np.abs(wdq).max() + np.abs(G[f"fl_{tag}_zbuf"]).max())
    assert np.max(np.abs(w - wdq)) <= tol


def test_rn_bf16_matches_torch():
    import torch
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(4096) * 10.0 ** rng.integers(-30, 30, 4096),
                        [0.0, 1.0, 1.00390625, 1.01171875, 3.38e38, 1e-40, -1e-40]]).astype(np.float32)
    got = O.rn_bf16(x.astype(np.float64))
    want = torch.from_numpy(x).to(torch.bfloat16).double().numpy()
    assert np.array_equal(got, want)


def test_forward_linearity_and_dispatch():
    rng = np.random.default_rng(1)
    N, K = 16, 256
    qw = rng.integers(-32768, 32767, (N // 4, K)).astype(np.int16)
    s = (0.004 + 0.012 * rng.random((O.scale_rows(K), N))).astype(np.float16)
    z = (-s.astype(np.float32) * rng.integers(0, 16, s.shape)).astype(np.float16)
    x = rng.standard_normal((3, K)).astype(np.float16)
    y = O.wq_linear_forward(x, qw, s, z)
    y2 = O.wq_linear_forward(2 * x.astype(np.float64), qw, s, z)
    assert np.allclose(y2, 2 * y, rtol=1e-12)
    # one-hot activation reads back a dequantised column exactly
    e = np.zeros((1, K)); e[0, 77] = 1.0
    assert np.array_equal(O.wq_linear_forward(e, qw, s, z)[0], O.dequant_weight(qw, s, z)[:, 77])
    assert O.dispatch_is_gemv((1, 7, K)) and not O.dispatch_is_gemv((1, 8, K))


def test_cpu_path_matches_numpy_oracle():
    """oracle/cpu_path.py (the timed pure-PyTorch CPU arm) == numpy oracle up to the operand rounding."""
    import torch
    from oracle import cpu_path
    rng = np.random.default_rng(3)
    N, K = 64, 1408            # 11 groups -> 16 scale rows (padding rows present)
    qw = rng.integers(-32768, 32767, (N // 4, K)).astype(np.int16)
    rows = O.scale_rows(K)
    s = np.zeros((rows, N), np.float16)
    s[: K // 128] = (0.004 + 0.012 * rng.random((K // 128, N))).astype(np.float16)
    z = (-s.astype(np.float32) * rng.integers(0, 16, s.shape)).astype(np.float16)
    x = rng.standard_normal((3, K)).astype(np.float16)
    tq, ts, tz, tx = (torch.from_numpy(a) for a in (qw, s, z, x))
    assert np.array_equal(cpu_path.unpack_intweight(tq).numpy(), O.unpack_intweight(qw))
    w_exact = O.unpack_intweight(qw).astype(np.float64) * np.repeat(s[: K // 128].T.astype(np.float64), 128, 1) \
        + np.repeat(z[: K // 128].T.astype(np.float64), 128, 1)
    assert np.allclose(cpu_path.dequant_weight(tq, ts, tz).double().numpy(), w_exact, rtol=1e-6, atol=1e-7)
    y = cpu_path.wq_linear_forward(tx, tq, ts, tz).double().numpy()
    y64 = O.wq_linear_forward(x, qw, s, z)
    assert np.linalg.norm(y - y64) / np.linalg.norm(y64) < 1e-3   # fp16 operand rounding of the kernels' contract
