"""The Hydragen attention operator -- the surface of the reference's ``hydragen/attention.py``
(same function names, argument meaning, return layouts and error behaviour), running on the
hand-written sm_100a kernels behind the C ABI.

Launches per call with L shared levels:  ONE tcgen05 prefix launch (L >= 2: the persistent kernel over all levels;
L = 1: the one-CTA-per-unit kernel) + 1 row-wise launch that does the suffix branch AND the combine -- versus, in the reference, L flash-attn launches +
L LSE transposes + cast + split-K + reduce + combine (Triton for 2 inputs, ~8 eager torch
launches otherwise; hydragen/attention.py:246-352, hydragen/flash.py:163-281).
"""

from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor

from . import _lib
from .flash import (
    causal_attention_tc,
    decode_attention_fused,
    flash_attention,
    flash_attention_seqlen,
    flash_attention_varlen,
    prefix_attention_grouped,
    prefix_attention_levels,
    prefix_attention_partials,
    suffix_attention_fused,
)

__all__ = [
    "combine_lse",
    "combine_lse_cuda",
    "combine_lse_triton",
    "combine_lse_torch",
    "hydragen_attention",
    "hydragen_attention_nopad",
    "hydragen_attention_decode",
    "flash_attention",
    "flash_attention_varlen",
    "flash_attention_seqlen",
]


def combine_lse_torch(outs: List[Tensor], lses: List[Tensor]):
    """Eager-torch statement of the combine, kept because the reference exposes it and its tests
    compare the kernel against it (hydragen/attention.py:21-43).  Not used by any hot path here."""
    o = torch.stack(outs)
    l = torch.stack(lses)
    w = (l - l.max(0).values[None]).exp()
    return ((o * w.unsqueeze(-1)).sum(0) / w.sum(0).unsqueeze(-1)).to(o.dtype)


def combine_lse_cuda(outs: List[Tensor], lses: List[Tensor], return_lse: bool = False):
    """n-way combine in one CUDA launch (any n <= 8, any head_dim, fp16/bf16/fp32).
    outs: list of [batch, seq_len, qheads, hdim]; lses: list of [batch, seq_len, qheads] fp32."""
    if len(outs) != len(lses) or len(outs) < 1:
        raise ValueError(f"need matching non-empty lists, got {len(outs)} outs and {len(lses)} lses")
    if len(outs) > _lib.HG_MAX_COMBINE:
        raise ValueError(f"at most {_lib.HG_MAX_COMBINE} partial results can be combined in one call")
    shape, dtype = outs[0].shape, outs[0].dtype
    for o, l in zip(outs, lses):
        if o.shape != shape or o.dtype != dtype:
            raise ValueError("all outs must share shape and dtype")
        if tuple(l.shape) != tuple(shape[:-1]):
            raise ValueError(f"lse shape {tuple(l.shape)} does not match out shape {tuple(shape)}")
        # the reference asserts contiguity (attention.py:122-126)
        assert o.is_contiguous(), "outs must be contiguous"
        assert l.is_contiguous(), "lses must be contiguous"
    lses = [l if l.dtype == torch.float32 else l.float() for l in lses]
    out = torch.empty_like(outs[0])
    lse_out = torch.empty(shape[:-1], device=out.device, dtype=torch.float32) if return_lse else None
    _lib.combine_lse(outs, lses, out, lse_out)
    return (out, lse_out) if return_lse else out


def combine_lse_triton(out1: Tensor, lse1: Tensor, out2: Tensor, lse2: Tensor):
    """Name kept for drop-in use (hydragen/attention.py:105-151); runs the CUDA kernel, not Triton."""
    return combine_lse_cuda([out1, out2], [lse1, lse2])


def combine_lse(outs: List[Tensor], lses: List[Tensor], enable_triton: bool = True):
    """Merge attention results using log-sum-exp metadata (hydragen/attention.py:154-174).

    ``enable_triton=True`` (the default, and what the operator uses) selects the CUDA kernel -- for
    ANY number of inputs, not just two; ``False`` selects the eager-torch statement, as in the
    reference."""
    if enable_triton:
        return combine_lse_cuda(outs, lses)
    return combine_lse_torch(outs, lses)


def hydragen_attention(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    shared_ks: List[Tensor],
    shared_vs: List[Tensor],
    shared_cu_seq_lens: List[Optional[Tensor]],
    shared_max_seq_lens: List[Optional[int]],
    use_varlens: List[bool],
    seq_lens: Optional[Tensor] = None,
):
    """Hydragen attention: attention decomposition + inter-sequence batching
    (hydragen/attention.py:177-354; argument meaning identical).

    q [batch, qlen, qheads, d]; k, v [batch, kvlen, kvheads, d] (unique per sequence);
    shared_ks[i] / shared_vs[i]: [sbatch, slen, kvheads, d] if not use_varlens[i], else packed
    [total_slen, kvheads, d] with shared_cu_seq_lens[i] (int32 [sbatch+1], starts with 0) and
    shared_max_seq_lens[i]; sbatch must divide batch and the batch is grouped contiguously by
    shared parent.  seq_lens: valid unique lengths per sequence (right padded) or None
    (no padding; the suffix branch is then causal, attention.py:344).
    """
    assert q.ndim == 4, f"{q.shape}"
    assert k.ndim == 4, f"{k.shape}"
    assert v.ndim == 4, f"{v.shape}"
    assert k.shape == v.shape
    assert len(shared_ks) == len(shared_vs) == len(shared_cu_seq_lens) == len(shared_max_seq_lens) == len(use_varlens)
    for sk, sv in zip(shared_ks, shared_vs):
        assert sk.shape == sv.shape, f"{sk.shape} {sv.shape}"

    b, nq, hq, d = q.shape
    n_groups = []
    for sk, scu, use_varlen in zip(shared_ks, shared_cu_seq_lens, use_varlens):
        n = scu.shape[0] - 1 if use_varlen else sk.shape[0]
        assert b % n == 0, f"{b} {n}"
        n_groups.append(n)
    # several shared levels: ONE persistent launch (the reference loops: one flash-attn call + LSE transpose per level);
    # a single level with few work items (TP ranks): split-KV partials, merged by the combine below
    early = k.shape[1] == 0 and len(shared_ks) == 1
    outs, lses = prefix_attention_levels(
        q, shared_ks, shared_vs, n_groups,
        [scu if uv else None for scu, uv in zip(shared_cu_seq_lens, use_varlens)],
        [smax if uv else None for smax, uv in zip(shared_max_seq_lens, use_varlens)],
        max_partials=1 if early else _lib.HG_MAX_COMBINE - 1)
    if early:
        return outs[0]  # attention.py:273-274, 330-331

    if k.shape[1] == 0:
        # >= 2 shared levels and no unique keys: undefined in the reference (flash-attn with
        # sk = 0); here simply the merge of the shared levels.
        return combine_lse_cuda(outs, lses)

    if seq_lens is None:
        # prefill of a chunk that attends to shared levels (hydragen/llama.py:513-521, 550-558): nq in the hundreds or
        # thousands -- the causal suffix branch (attention.py:344) is dense work for the tensor cores, then one combine
        res = causal_attention_tc(q, k, v)
        if res is not None:
            return combine_lse_cuda(outs + [res[0]], lses + [res[1]])
    # decode / short chunks: suffix branch + (L+1)-way combine in one CUDA-core launch
    out, _ = suffix_attention_fused(q, k, v, seq_lens, causal=seq_lens is None, partial_outs=outs, partial_lses=lses)
    return out


def hydragen_attention_decode(
    q: Tensor,
    k_new: Tensor,
    v_new: Tensor,
    positions: Tensor,
    k_cache: Tensor,
    v_cache: Tensor,
    shared_ks: List[Tensor],
    shared_vs: List[Tensor],
    shared_cu_seq_lens: Optional[List[Optional[Tensor]]] = None,
    shared_max_seq_lens: Optional[List[Optional[int]]] = None,
    use_varlens: Optional[List[bool]] = None,
):
    """A whole decode step of Hydragen attention for one layer: what the reference's DECODE branch does
    with ``update_per_completion_kvs`` followed by ``hydragen_attention(..., seq_lens=pos + 1)``
    (hydragen/llama.py:564-587), in TWO launches: one persistent tcgen05 prefix launch over every shared
    level, then ONE launch that appends the new token's K/V at ``positions[b]``, runs the suffix branch over
    the sequence's ``positions[b] + 1`` own keys and merges every partial result.  The caches are updated
    in place; returns ``out [b, 1, hq, d]``."""
    n = len(shared_ks)
    shared_cu_seq_lens = shared_cu_seq_lens or [None] * n
    shared_max_seq_lens = shared_max_seq_lens or [None] * n
    use_varlens = use_varlens or [False] * n
    assert q.ndim == 4 and q.shape[1] == 1, f"{q.shape}"
    assert len(shared_vs) == n and len(shared_cu_seq_lens) == n and len(shared_max_seq_lens) == n and len(use_varlens) == n
    b = q.shape[0]
    n_groups = []
    for sk, sv, scu, use_varlen in zip(shared_ks, shared_vs, shared_cu_seq_lens, use_varlens):
        assert sk.shape == sv.shape, f"{sk.shape} {sv.shape}"
        ng = scu.shape[0] - 1 if use_varlen else sk.shape[0]
        assert b % ng == 0, f"{b} {ng}"
        n_groups.append(ng)
    outs, lses = prefix_attention_levels(
        q, shared_ks, shared_vs, n_groups,
        [scu if uv else None for scu, uv in zip(shared_cu_seq_lens, use_varlens)],
        [smax if uv else None for smax, uv in zip(shared_max_seq_lens, use_varlens)],
        max_partials=_lib.HG_MAX_COMBINE)
    out, _ = decode_attention_fused(q, k_new, v_new, positions, k_cache, v_cache, outs, lses)
    return out


def hydragen_attention_nopad(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    shared_ks: List[Tensor],
    shared_vs: List[Tensor],
    seq_len: Optional[Tensor] = None,
):
    """hydragen/attention.py:357-392: every shared level is [sbatch, slen, kvheads, d], no padding."""
    n = len(shared_ks)
    return hydragen_attention(
        q, k, v,
        shared_ks=shared_ks, shared_vs=shared_vs,
        shared_cu_seq_lens=[None] * n, shared_max_seq_lens=[None] * n, use_varlens=[False] * n,
        seq_lens=seq_len,
    )
