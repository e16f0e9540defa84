"""CPU oracle for the ops adjacent to the W4A16 linears inside a decoder layer.

TEST INFRASTRUCTURE ONLY -- same rules as ``oracle/w4a16_oracle.py``: nothing under ``llm_awq_b200/`` may import
this file.

Parity status: the reference has no tests or golden vectors for these ops and its implementations are CUDA-only
(no CPU path to import here), so the restatements below are pinned on the GPU box against the reference's own
kernels recompiled for sm_100a (``oracle/_ref``: ``layernorm_forward_cuda``) and, for the fused MLP front half,
against the composition the reference executes (two engine calls + ``F.silu`` + multiply,
``tinychat/modules/fused_mlp.py:36-83``) evaluated with torch on the same device.

Each function cites the reference file:line it restates.
"""
from __future__ import annotations

import numpy as np

from . import w4a16_oracle as O

HALF_CLAMP = 65504.0 - 1000.0   # awq/kernels/csrc/layernorm/reduction.cuh:81 (HALF_FLT_MAX - 1000)


def rmsnorm(x: np.ndarray, gamma: np.ndarray, eps: float, dtype: str = "f16", rounded: bool = True) -> np.ndarray:
    """``generalT5LayerNorm`` (awq/kernels/csrc/layernorm/layernorm.cu:38-64): per row,
    ``var = sum x^2`` (:50-53), ``r = rsqrt(var / n + eps)`` (:57), ``y = clamp_T((x * r) * gamma)`` (:62; the clamp is
    fp16-only, reduction.cuh:72-82).  float64 statistics; one rounding to ``dtype`` when ``rounded``."""
    x = np.asarray(x, dtype=np.float64)
    g = np.asarray(gamma, dtype=np.float64)
    n = x.shape[-1]
    r = 1.0 / np.sqrt((x * x).sum(axis=-1, keepdims=True) / n + eps)
    y = (x * r) * g
    if not rounded:
        return y
    if dtype == "f16":
        y = np.clip(y, -HALF_CLAMP, HALF_CLAMP)
    return O.rounder(dtype)(y)


def silu(x: np.ndarray) -> np.ndarray:
    """``F.silu`` (tinychat/modules/fused_mlp.py:50,78): x * sigmoid(x), float64."""
    x = np.asarray(x, dtype=np.float64)
    return x / (1.0 + np.exp(-x))


def mlp_front(x, gate, up, dtype: str = "f16", rows=None) -> np.ndarray:
    """``QuantLlamaMLP.our_llama_mlp`` (tinychat/modules/fused_mlp.py:36-83) with the zeros tensors AS GIVEN
    (the caller applies the ``- 8 * scales`` of the GEMM branch, :69,76, when it wants that branch):
        gate_out = rn_T(x @ w~_gate^T)   (:40-49 / :65-71)      up_out = rn_T(x @ w~_up^T)   (:51-60 / :72-77)
        c = rn_T( rn_T(silu(gate_out)) * up_out )               (:50,78,80)
    ``gate`` / ``up`` are (qweight, scales, szeros) triples.  Every intermediate is rounded to ``dtype`` exactly
    where the reference materialises a tensor."""
    rn = O.rounder(dtype)
    g = rn(O.wq_linear_forward(x, *gate, dtype=dtype, rows=rows))
    u = rn(O.wq_linear_forward(x, *up, dtype=dtype, rows=rows))
    return rn(rn(silu(g)) * u)
