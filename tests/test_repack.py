"""v1 -> v2 repacking and tensor-parallel state-dict loading on CPU, pinned to vectors produced by
the reference's own tinychat/offline-weight-repacker.py (tests/golden/make_golden.py::repacker_golden)."""
import os

import numpy as np
import pytest
import torch

import llm_awq_b200 as P
from llm_awq_b200 import repack
from oracle import w4a16_oracle as O
from awq_testutil import gen_layer, np64

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_repacker.npz"))


@pytest.mark.parametrize("idx", range(4))
def test_v1_to_v2_matches_the_reference_repacker_bit_for_bit(idx):
    qw = torch.from_numpy(G[f"r{idx}_qweight_v1"])
    sc = torch.from_numpy(G[f"r{idx}_scales_v1"]).half()
    qz = torch.from_numpy(G[f"r{idx}_qzeros_v1"])
    assert np.array_equal(repack.qweight_unpack_v1(qw).numpy().astype(np.uint8), G[f"r{idx}_unpacked"])
    assert np.array_equal(repack.qweight_pack_v1_to_v2(qw).numpy(), G[f"r{idx}_qweight_v2"])
    assert np.array_equal(repack.multiply_scale_qzero_negative(sc, qz).float().numpy(), G[f"r{idx}_scaled_zeros"])
    # and the oracle's closed-form unpacker reads the v2 tensor back to the v1 nibbles
    assert np.array_equal(O.unpack_intweight(G[f"r{idx}_qweight_v2"]), G[f"r{idx}_unpacked"])


def test_repack_state_dict_keys_and_layouts():
    N, K = 64, 1024
    sd = {"model.layers.0.mlp.up_proj.qweight": torch.from_numpy(G["r2_qweight_v1"]),
          "model.layers.0.mlp.up_proj.scales": torch.from_numpy(G["r2_scales_v1"]).half(),
          "model.layers.0.mlp.up_proj.qzeros": torch.from_numpy(G["r2_qzeros_v1"]),
          "model.norm.weight": torch.ones(8)}
    out = repack.repack_state_dict_v1_to_v2(sd)
    assert set(out) == {"model.layers.0.mlp.up_proj.qweight", "model.layers.0.mlp.up_proj.scales",
                        "model.layers.0.mlp.up_proj.scaled_zeros", "model.norm.weight"}
    assert out["model.layers.0.mlp.up_proj.qweight"].shape == (N // 4, K) and out["model.layers.0.mlp.up_proj.qweight"].dtype == torch.int16
    assert out["model.layers.0.mlp.up_proj.scales"].shape == (K // 128, N)                 # transposed, not padded (:133)
    assert np.array_equal(out["model.layers.0.mlp.up_proj.scaled_zeros"].float().numpy(), G["r2_scaled_zeros"].T)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_state_dict_reassembles_to_the_full_layer(world, tmp_path):
    """Column shards concatenate, row shards add: y_full == sum_r y_row_r on the oracle; fused QKV shards by head block."""
    dtype = torch.float16
    H, I, KV = 512, 1024, 256
    names = {"self_attn.qkv_proj": (H + 2 * KV, H), "self_attn.o_proj": (H, H), "mlp.gate_proj": (I, H),
             "mlp.up_proj": (I, H), "mlp.down_proj": (H, I)}
    sd = {}
    for i, (nm, (N, K)) in enumerate(names.items()):
        qw, s, z = gen_layer(N, K, dtype, seed=i)
        sd[f"layers.0.{nm}.qweight"], sd[f"layers.0.{nm}.scales"], sd[f"layers.0.{nm}.scaled_zeros"] = qw, s, z
    sd["layers.0.mlp.down_proj.bias"] = torch.randn(H).half()
    sd["norm.weight"] = torch.ones(H)
    path = str(tmp_path / "model-v2.pt")
    torch.save(sd, path)
    shards = [repack.load_quant_sharded(path, r, world, fused_qkv=(H, KV)) for r in range(world)]
    assert all(torch.equal(s["norm.weight"], sd["norm.weight"]) for s in shards)
    x = torch.randn(3, H).half()
    f = lambda d, nm, xx: O.wq_linear_forward(np64(xx), d[f"layers.0.{nm}.qweight"].numpy(), np64(d[f"layers.0.{nm}.scales"]),
                                              np64(d[f"layers.0.{nm}.scaled_zeros"]))
    # column-parallel: concatenation over ranks == full
    for nm in ("mlp.gate_proj", "mlp.up_proj"):
        assert np.array_equal(np.concatenate([f(s, nm, x) for s in shards], axis=1), f(sd, nm, x))
    # fused QKV: per rank [q_r; k_r; v_r]
    full = f(sd, "self_attn.qkv_proj", x)
    q, k, v = full[:, :H], full[:, H:H + KV], full[:, H + KV:]
    for r, s in enumerate(shards):
        got = f(s, "self_attn.qkv_proj", x)
        hq, hk = H // world, KV // world
        want = np.concatenate([q[:, r * hq:(r + 1) * hq], k[:, r * hk:(r + 1) * hk], v[:, r * hk:(r + 1) * hk]], axis=1)
        assert np.array_equal(got, want)
    # row-parallel: partial sums over ranks == full (float64 oracle: exact up to summation order)
    xi = torch.randn(3, I).half()
    part = sum(f(s, "mlp.down_proj", xi[:, r * (I // world):(r + 1) * (I // world)]) for r, s in enumerate(shards))
    assert np.allclose(part, f(sd, "mlp.down_proj", xi), rtol=1e-12, atol=1e-9)
    assert torch.equal(shards[0]["layers.0.mlp.down_proj.bias"], sd["layers.0.mlp.down_proj.bias"])
    assert torch.count_nonzero(shards[1]["layers.0.mlp.down_proj.bias"]) == 0          # the bias is added once
    # scale rows are re-padded to a multiple of 8 (WQLinear's buffer shape for the shard, qmodule.py:109-130)
    assert shards[0]["layers.0.mlp.down_proj.scales"].shape[0] == O.scale_rows(I // world)


def test_load_quant_sharded_repacks_v1_checkpoints(tmp_path):
    sd = {"l.up_proj.qweight": torch.from_numpy(G["r3_qweight_v1"]), "l.up_proj.scales": torch.from_numpy(G["r3_scales_v1"]).half(),
          "l.up_proj.qzeros": torch.from_numpy(G["r3_qzeros_v1"])}
    path = str(tmp_path / "model-v1.pt")
    torch.save(sd, path)
    out = repack.load_quant_sharded(path, 0, 1)
    assert np.array_equal(out["l.up_proj.qweight"].numpy(), G["r3_qweight_v2"]) and "l.up_proj.scaled_zeros" in out
