"""On-disk formats adjacent to the W4A16 path: v1 -> v2 checkpoint repacking and tensor-parallel
loading of `*-v2.pt` state dicts.  Host-side, vectorised torch (works on CPU and on the GPU; no Python loop over
columns or groups).

v1 (the first llm-awq / AutoAWQ "GEMM" format; reference: tinychat/offline-weight-repacker.py:8-79)
    qweight  int32 [N, K / 8]      nibble j (bits 4j..4j+3) of word c is the weight of input channel 8c + j
    scales   T     [N, K / G]
    qzeros   int32 [N, K / G / 8]  nibble j of word c is the zero point of group 8c + j
v2 (what WQLinear holds; awq/quantize/qmodule.py:98-137)
    qweight  int16 [N / 4, K]      pack_intweight layout (qmodule.py:26-65)
    scales   T     [ceil8(K / G), N]           (transposed; the repacker does NOT pad: offline-weight-repacker.py:133)
    scaled_zeros = -(scales * zero + zp_shift * scales)  with zp_shift = -8 for v1 checkpoints (:66-76)
"""
from collections import OrderedDict

import torch

from .qmodule import pack_intweight
from . import tp as _tp


def qweight_unpack_v1(qweight: torch.Tensor) -> torch.Tensor:
    """int32 [N, K / 8] -> int32 [N, K] of nibbles (offline-weight-repacker.py:8-19, vectorised)."""
    assert qweight.dtype == torch.int32
    shifts = torch.arange(0, 32, 4, device=qweight.device, dtype=torch.int32)
    return ((qweight.unsqueeze(-1) >> shifts) & 0xF).reshape(qweight.shape[0], -1)


def qweight_pack_v1_to_v2(qweight: torch.Tensor, interleave: int = 4, kstride: int = 64) -> torch.Tensor:
    """offline-weight-repacker.py:76-79 (qweight_unpack + packing_v2_from_unpacked, :22-63)."""
    assert interleave == 4 and kstride == 64
    return pack_intweight(qweight_unpack_v1(qweight))


def multiply_scale_qzero_negative(scales: torch.Tensor, qzeros: torch.Tensor, zp_shift: int = -8) -> torch.Tensor:
    """scaled_zeros [N, K / G] = -(scales * zero + zp_shift * scales) (offline-weight-repacker.py:66-76), evaluated in
    the scales' dtype with the reference's operation order."""
    shifts = torch.arange(0, 32, 4, device=qzeros.device, dtype=torch.int32)
    zeros = ((qzeros.unsqueeze(-1) >> shifts) & 0xF).reshape(qzeros.shape[0], -1)[:, : scales.shape[1]]
    scaled = scales * zeros.to(scales.dtype)
    return -(scaled + (zp_shift * scales))


def repack_state_dict_v1_to_v2(model_dict, zp_shift: int = -8):
    """The whole of `offline_repacker` (offline-weight-repacker.py:107-160): qweight -> v2 packing, scales transposed,
    qzeros replaced by scaled_zeros (transposed), everything else copied."""
    out = OrderedDict()
    for key, param in model_dict.items():
        if "qweight" in key:
            out[key] = qweight_pack_v1_to_v2(param)
        elif "scales" in key:
            out[key] = param.transpose(1, 0).contiguous()
            zkey = key.replace("scales", "qzeros")
            out[key.replace("scales", "scaled_zeros")] = multiply_scale_qzero_negative(
                param, model_dict[zkey], zp_shift).transpose(1, 0).contiguous()
        elif "qzeros" in key:
            continue
        else:
            out[key] = param
    return out


# Llama-family projection names -> how they shard
COLUMN_PARALLEL = ("q_proj", "k_proj", "v_proj", "gate_proj", "up_proj")
ROW_PARALLEL = ("o_proj", "down_proj")


def shard_state_dict_v2(sd, rank: int, world: int, group_size: int = 128, fused_qkv=None):
    """Rank `rank`'s tensors of a v2 state dict under Megatron-style tensor parallelism: column-parallel projections
    keep out_features / world rows of qweight (contiguous row slices of the packed tensor) and the matching columns of
    scales / scaled_zeros / bias; row-parallel ones keep in_features / world int16 columns and the matching group rows
    (re-padded to a multiple of 8), bias on rank 0 only.  `fused_qkv` = (q_out, kv_out) shards a fused `qkv_proj`
    (tinychat/modules/fused_attn.py:566-594) by head block.  Other tensors are replicated."""
    out = OrderedDict()
    done = set()
    for key in sd:
        if key in done or not key.endswith("qweight"):
            continue
        prefix = key[: -len("qweight")]
        qw, sc, sz = sd[key], sd[prefix + "scales"], sd[prefix + "scaled_zeros"]
        bias = sd.get(prefix + "bias")
        name = prefix.rstrip(".").rsplit(".", 1)[-1]
        if name == "qkv_proj" and fused_qkv is not None:
            q, s, z, b = _tp.shard_fused_qkv(qw, sc, sz, bias, fused_qkv[0], fused_qkv[1], rank, world)
        elif name in COLUMN_PARALLEL or name == "qkv_proj":
            q, s, z, b = _tp.shard_column(qw, sc, sz, bias, rank, world)
        elif name in ROW_PARALLEL:
            q, s, z = _tp.shard_row(qw, sc, sz, rank, world, group_size)
            b = bias if (bias is not None and rank == 0) else (None if bias is None else torch.zeros_like(bias))
        else:
            continue
        out[prefix + "qweight"], out[prefix + "scales"], out[prefix + "scaled_zeros"] = q, s, z
        done.update({prefix + "qweight", prefix + "scales", prefix + "scaled_zeros"})
        if bias is not None:
            out[prefix + "bias"] = b
            done.add(prefix + "bias")
    for key, v in sd.items():
        if key not in done:
            out[key] = v
    return out


def load_quant_sharded(checkpoint: str, rank: int, world: int, group_size: int = 128, fused_qkv=None, map_location="cpu"):
    """torch.load of a `*-v2.pt` (or v1: detected by int32 qweights and repacked first) state dict, sharded for `rank`
    (the loading half of tinychat/utils/load_quant.py:145-171, without the model surgery)."""
    sd = torch.load(checkpoint, map_location=map_location)
    if any(k.endswith("qweight") and v.dtype == torch.int32 for k, v in sd.items()):
        sd = repack_state_dict_v1_to_v2(sd)
    return sd if world == 1 else shard_state_dict_v2(sd, rank, world, group_size, fused_qkv)
