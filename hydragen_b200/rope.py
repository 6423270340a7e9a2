"""RoPE of the new q / k rows in ONE launch -- the step right before the attention hot path
(SURVEY.md 8f, row N2).

Same name and argument meaning as the call the reference makes (hydragen/llama.py:494-501 ->
``transformers.models.llama.modeling_llama.apply_rotary_pos_emb`` of 4.37.2, pinned upstream, not in tree):
``cos`` / ``sin`` are the FULL cached tables ``[max_pos, d]`` in the activation dtype
(``HydragenLlamaRotaryEmbedding.forward``, hydragen/llama.py:47-55) and ``position_ids [b, s]`` selects a
row per token; the gather, ``rotate_half`` and the five elementwise operations per tensor are one kernel
(``hg_rope_qk``, csrc/rope.cu) whose output is bit-identical to the eager evaluation.  No CPU fallback.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib


def _row_stride(t: Tensor) -> Optional[int]:
    """Uniform row stride of a [b, s, h, d] tensor whose (b, s) axes collapse and whose heads are dense."""
    b, s, h, d = t.shape
    if t.stride(3) != 1 or (h > 1 and t.stride(2) != d):
        return None
    if b * s <= 1:
        return max(h * d, 1)
    if s == 1:
        rs = t.stride(0)
    elif b == 1:
        rs = t.stride(1)
    else:
        if t.stride(0) != s * t.stride(1):
            return None
        rs = t.stride(1)
    return rs if rs >= h * d else None


def apply_rotary_pos_emb(q: Tensor, k: Tensor, cos: Tensor, sin: Tensor, position_ids: Tensor, unsqueeze_dim: int = 2,
                         inplace: bool = False) -> Tuple[Tensor, Tensor]:
    """q ``[b, s, hq, d]``, k ``[b, s, hkv, d]`` (views of a fused projection output are fine), cos / sin
    ``[max_pos, d]`` of q's dtype, position_ids ``[b, s]`` (int32 / int64) absolute positions.  Returns the
    rotated (q, k); with ``inplace=True`` the inputs are overwritten and returned."""
    if unsqueeze_dim != 2:
        raise NotImplementedError("the Hydragen path keeps heads on axis 2 (hydragen/llama.py:500)")
    if q.ndim != 4 or k.ndim != 4 or q.shape[:2] != k.shape[:2] or q.shape[3] != k.shape[3]:
        raise ValueError(f"expected q [b, s, hq, d] and k [b, s, hkv, d], got {tuple(q.shape)} {tuple(k.shape)}")
    b, s, hq, d = q.shape
    hkv = k.shape[2]
    if q.dtype != k.dtype or cos.dtype != q.dtype or sin.dtype != q.dtype:
        raise ValueError(f"q/k/cos/sin dtypes differ: {q.dtype} {k.dtype} {cos.dtype} {sin.dtype}")
    if cos.shape != sin.shape or cos.ndim != 2 or cos.shape[1] != d:
        raise ValueError(f"cos / sin must be the [max_pos, {d}] tables, got {tuple(cos.shape)} {tuple(sin.shape)}")
    if position_ids.numel() != b * s:
        raise ValueError(f"position_ids must hold one position per token ({b} x {s}), got {tuple(position_ids.shape)}")
    q_rs, k_rs = _row_stride(q), _row_stride(k)
    if q_rs is None:
        q, inplace_q = q.contiguous(), True
        q_rs = hq * d
    else:
        inplace_q = inplace
    if k_rs is None:
        k, inplace_k = k.contiguous(), True
        k_rs = hkv * d
    else:
        inplace_k = inplace
    q_out = q if inplace_q else torch.empty((b, s, hq, d), device=q.device, dtype=q.dtype)
    k_out = k if inplace_k else torch.empty((b, s, hkv, d), device=k.device, dtype=k.dtype)
    _lib.rope_qk(q, k, q_out, k_out, cos.contiguous(), sin.contiguous(), position_ids.reshape(-1).contiguous(), b * s, hq, hkv, d,
                 q_rs, k_rs, q_rs if inplace_q else hq * d, k_rs if inplace_k else hkv * d)
    return q_out, k_out
