"""The "pure-PyTorch dequant" CPU statement of the W4A16 linear path, multi-threaded.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/w4a16_oracle.py header): used by
``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference`` as the timed CPU arm,
and by ``tests/`` (checked against the numpy oracle).  Nothing under ``llm_awq_b200/``
imports it.  Parity status: pinned by us, through the numpy oracle (itself pinned to the
reference's Python packer / quantiser outputs in tests/golden/).

The reference has NO CPU forward for WQLinear (its kernels are CUDA-only).
Its only PyTorch statement of the arithmetic is the fake-quant formula
``w = (q - z) * s`` (awq/quantize/quantizer.py:93-95) with the packed-buffer conventions of
awq/quantize/qmodule.py:157-197 (``scaled_zeros = -s * z``).  This file restates exactly
that with torch tensor ops on the host:  unpack qweight (closed form of
qmodule.py:26-65) -> W = q * S + Z per 128-group (fp32) -> y = x @ W^T (fp32).
"""
from __future__ import annotations

import torch


def unpack_intweight(qweight: torch.Tensor) -> torch.Tensor:
    """int16 [N/4, K] -> uint8 [N, K] nibbles (inverse of pack_intweight, qmodule.py:26-65).

    P[r, 64*tile + 16*row + 8*blk + 2*u + hi] nibble j  ==  q[4r + row, 64*tile + 32*blk + 8*j + 2*u + hi].
    """
    R, K = qweight.shape
    p = (qweight.to(torch.int32) & 0xFFFF).view(R, K // 64, 4, 2, 4, 2)          # r, tile, row, blk, u, hi
    shifts = torch.arange(0, 16, 4, dtype=torch.int32)
    nib = (p.unsqueeze(-1) >> shifts) & 0xF                                        # ..., j
    q = nib.permute(0, 2, 1, 3, 6, 4, 5)                                           # r, row, tile, blk, j, u, hi
    return q.reshape(R * 4, K).to(torch.uint8)


def dequant_weight(qweight, scales, szeros, group_size: int = 128) -> torch.Tensor:
    """fp32 [N, K]:  q * S[k // G, n] + Z[k // G, n]   (no rounding to the activation dtype: this is
    the plain PyTorch formula, not the kernels' operand rounding -- the numpy oracle models that)."""
    q = unpack_intweight(qweight).to(torch.float32)
    N, K = q.shape
    ng = K // group_size
    S = scales[:ng].to(torch.float32).t().contiguous()                              # [N, K/G]
    Z = szeros[:ng].to(torch.float32).t().contiguous()
    w = q.view(N, ng, group_size) * S.unsqueeze(-1) + Z.unsqueeze(-1)
    return w.view(N, K)


def wq_linear_forward(x, qweight, scales, szeros, bias=None, group_size: int = 128) -> torch.Tensor:
    """y = x @ dequant(W)^T (+ bias) in fp32; the dequantisation is redone on every call, as in
    weight-only-quantised inference (WQLinear.forward, qmodule.py:201-224)."""
    w = dequant_weight(qweight, scales, szeros, group_size)
    y = x.reshape(-1, x.shape[-1]).to(torch.float32) @ w.t()
    if bias is not None:
        y = y + bias.to(torch.float32)
    return y.reshape(*x.shape[:-1], w.shape[0])
