"""Attention primitives returning ``(out, lse)`` -- the surface of the reference's
``hydragen/flash.py`` (same names, argument meaning and return layouts), implemented by the
hand-written sm_100a kernels behind the C ABI.  No flash-attn, no Triton, no CPU fallback.

=====================================  =======================================================
reference (hydragen/flash.py)          here
=====================================  =======================================================
``flash_attention`` :284-306           tcgen05 prefix kernel (16-bit, d in {64,128}; causal = its masked
  (flash-attn _flash_attn_forward)     instantiation, sk >= sq); row-wise kernel otherwise (fp32, tiny chunks)
``flash_attention_varlen`` :309-351    tcgen05 prefix kernel with a device ``cu_seqlens_k`` table
``flash_attention_seqlen`` :163-281    row-wise kernel (one launch instead of cast + split-K +
  (+ Triton kernels, pick_split_k)     reduce); warps-per-sequence chosen from the cache length
=====================================  =======================================================

LSE convention (all): fp32 natural log of sum exp(q.k / sqrt(d)).  ``flash_attention`` /
``flash_attention_varlen`` return it shaped ``[b, h, sq]`` like flash-attn v2.3.6 does -- as a
permuted VIEW of the ``[b, sq, h]`` buffer the kernels write, so the reference's
``rearrange(...).contiguous()`` on it is a no-op instead of a transpose pass.
"""

from __future__ import annotations

import os
from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

_TC_DTYPES = (torch.float16, torch.bfloat16)
_TC_HEAD_DIMS = (64, 128)


def _causal_backend() -> str:
    """'auto' (default: tcgen05 when the shape allows), 'tcgen05' or 'rowwise' -- test hook, read per call."""
    return os.environ.get("HYDRAGEN_B200_CAUSAL_BACKEND", "auto")


def _prefix_backend() -> str:
    """'auto' (default), 'tcgen05' or 'rowwise' -- test hook, read per call."""
    return os.environ.get("HYDRAGEN_B200_PREFIX_BACKEND", "auto")


def _prefix_split() -> bool:
    """HYDRAGEN_B200_PREFIX_SPLIT=0: never cut a (group, tile, head) unit between CTAs (test hook, read per call)."""
    return os.environ.get("HYDRAGEN_B200_PREFIX_SPLIT", "1") != "0"


def _rows_view(t: Tensor) -> Optional[int]:
    """Row stride (elements) if the (b, s) axes of a [b, s, h, d] tensor collapse into one row axis
    with a uniform stride and each row is a dense [h, d] block; None otherwise."""
    b, s, h, d = t.shape
    if t.stride(3) != 1 or (h > 1 and t.stride(2) != d):
        return None
    if s == 1:
        rs = t.stride(0) if b > 1 else h * d
    elif b == 1:
        rs = t.stride(1)
    else:
        if t.stride(0) != s * t.stride(1):
            return None
        rs = t.stride(1)
    return rs if rs >= h * d else None


def _inner_ok(t: Tensor) -> bool:
    return t.stride(-1) == 1


def _check_qkv(q: Tensor, k: Tensor, v: Tensor):
    if q.ndim != 4 or k.ndim != 4 or v.ndim != 4:
        raise ValueError(f"expected 4-d q/k/v, got {tuple(q.shape)} {tuple(k.shape)} {tuple(v.shape)}")
    if k.shape != v.shape:
        raise ValueError(f"k/v shape mismatch {tuple(k.shape)} {tuple(v.shape)}")
    if q.shape[-1] != k.shape[-1]:
        raise ValueError(f"Keys have head dim {k.shape[-1]} but queries have head dim {q.shape[-1]}")
    if q.dtype != k.dtype or q.dtype != v.dtype:
        raise ValueError("q/k/v dtypes differ")
    if q.shape[2] % k.shape[2] != 0:
        raise ValueError(f"qheads {q.shape[2]} must be a multiple of kvheads {k.shape[2]}")


def prefix_attention_grouped(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    n_groups: int,
    cu_seqlens_k: Optional[Tensor] = None,
    max_seqlen_k: Optional[int] = None,
) -> Tuple[Tensor, Tensor]:
    """The prefix branch of ONE shared level as one result: ``q [b, nq, hq, d]`` with the batch grouped contiguously by
    shared parent (``b % n_groups == 0``), ``k, v`` either ``[n_groups, L, hkv, d]`` or, with ``cu_seqlens_k`` (int32
    ``[n_groups + 1]`` on device), packed ``[total, hkv, d]``.  Returns ``out [b, nq, hq, d]`` and ``lse [b, nq, hq]``
    (fp32) -- already in the layout the combine consumes (hydragen/attention.py:276-280, 333-338 need no transpose)."""
    outs, lses = prefix_attention_partials(q, k, v, n_groups, cu_seqlens_k, max_seqlen_k, max_splits=1)
    return outs[0], lses[0]


def _level_views(q, k, v, ng, cu, mx, hkv, d):
    """(k, v, n_k_rows, k_len, kv_row_stride, max_k) of one shared level for the tensor-core kernels."""
    if cu is not None:
        if not (k.stride(2) == 1 and k.stride(1) == d and v.stride() == k.stride()):
            k, v = k.contiguous(), v.contiguous()
        return k, v, k.shape[0], 0, k.stride(0), (int(mx) if mx is not None else k.shape[0])
    kv_rs = _rows_view(k)
    if kv_rs is None or _rows_view(v) != kv_rs:
        k, v = k.contiguous(), v.contiguous()
        kv_rs = hkv * d
    return k, v, k.shape[0] * k.shape[1], k.shape[1], kv_rs, k.shape[1]


def _check_level(q, k, v, ng, cu, i=0):
    b = q.shape[0]
    if ng < 1 or b % ng != 0:
        raise ValueError(f"batch {b} is not a multiple of the number of shared sequences {ng}")
    if k.shape != v.shape or k.shape[-1] != q.shape[-1] or k.dtype != q.dtype or v.dtype != q.dtype:
        raise ValueError(f"shared K/V of level {i}: shapes {tuple(k.shape)} / {tuple(v.shape)}, dtypes {k.dtype} / {v.dtype}")
    if cu is not None:
        if k.ndim != 3:
            raise ValueError("varlen shared K/V must be [total, kvheads, d]")
        if cu.dtype != torch.int32 or cu.shape[0] != ng + 1:
            raise ValueError("cu_seqlens_k must be int32 of n_groups + 1 entries")
    elif k.ndim != 4 or k.shape[0] != ng:
        raise ValueError(f"shared K/V must be [n_groups, L, kvheads, d], got {tuple(k.shape)}")


def prefix_attention_partials(q, k, v, n_groups, cu_seqlens_k=None, max_seqlen_k=None, max_splits: int = 1):
    """The prefix branch of ONE shared level as a list of partial results: when the launch has too few (group, tile,
    head) work items to fill the GPU -- the head-parallel ranks of a tensor-parallel run -- the keys are cut into up to
    ``max_splits`` ranges (split-KV) and one ``(out, lse)`` pair per range is returned, to be merged by the combine that
    follows anyway.  Returns (list of out, list of lse)."""
    b, nq, hq, d = q.shape
    _check_level(q, k, v, n_groups, cu_seqlens_k)
    hkv = k.shape[-2]
    sm_scale = d**-0.5  # hydragen/flash.py:293
    backend = _prefix_backend()
    use_tc = q.dtype in _TC_DTYPES and d in _TC_HEAD_DIMS and backend != "rowwise"
    if backend == "tcgen05" and not use_tc:
        raise ValueError(f"tcgen05 prefix kernel does not take dtype {q.dtype} / head_dim {d}")
    varlen = cu_seqlens_k is not None
    splits = 1
    if use_tc and max_splits > 1:
        k_max = int(max_seqlen_k) if (varlen and max_seqlen_k is not None) else (k.shape[0] if varlen else k.shape[1])
        splits = _lib.prefix_suggest_splits(q.device, n_groups, (b // n_groups) * nq, hq, k_max, max_splits)
    out = torch.empty((splits, b, nq, hq, d), device=q.device, dtype=q.dtype)
    lse = torch.empty((splits, b, nq, hq), device=q.device, dtype=torch.float32)
    if use_tc:
        q_rs = _rows_view(q)
        if q_rs is None:
            q = q.contiguous()
            q_rs = hq * d
        k, v, n_k_rows, k_len, kv_rs, max_k = _level_views(q, k, v, n_groups, cu_seqlens_k, max_seqlen_k, hkv, d)
        _lib.prefix_attn_fwd(q, k, v, out, lse, n_groups, (b // n_groups) * nq, n_k_rows, k_len, cu_seqlens_k, max_k,
                             hq, hkv, d, q_rs, kv_rs, sm_scale, kv_splits=splits)
    else:
        # CUDA-core path (fp32, other head dims): every sequence walks its parent's keys.
        if not _inner_ok(q):
            q = q.contiguous()
        if not _inner_ok(k):
            k = k.contiguous()
        if not _inner_ok(v) or v.stride() != k.stride():
            v = v.contiguous()
            k = k.contiguous()
        if varlen:
            strides = (0, k.stride(0), k.stride(1))
            lk = int(max_seqlen_k) if max_seqlen_k is not None else k.shape[0]
        else:
            strides = (k.stride(0), k.stride(1), k.stride(2))
            lk = k.shape[1]
        _lib.rowwise_attn_fwd(q, k, v, None, cu_seqlens_k, b // n_groups, False, out[0], lse[0], lk, strides, [], [], sm_scale)
    return [out[i] for i in range(splits)], [lse[i] for i in range(splits)]


def prefix_attention_levels(
    q: Tensor,
    shared_ks: Sequence[Tensor],
    shared_vs: Sequence[Tensor],
    n_groups: Sequence[int],
    cu_seqlens: Sequence[Optional[Tensor]],
    max_seqlens: Sequence[Optional[int]],
    max_partials: int = 1,
):
    """The prefix branch of EVERY shared level of a hierarchy (the loop of hydragen/attention.py:250-341): returns
    ``(outs, lses)`` lists of ``[b, nq, hq, d]`` / ``[b, nq, hq]`` partial results for the combine.

    * two or more levels (16-bit, d in {64, 128}): ONE persistent tcgen05 launch over all of them
      (hg_prefix_attn_grouped_fwd), one partial per level;
    * one level: the one-CTA-per-unit tcgen05 kernel; with ``max_partials > 1`` and few work items (the ranks of a
      tensor-parallel run) its keys are cut into up to that many ranges, one partial each;
    * fp32 inputs and other head dims: level by level on the CUDA-core kernel."""
    b, nq, hq, d = q.shape
    n_levels = len(shared_ks)
    if not (len(shared_vs) == len(n_groups) == len(cu_seqlens) == len(max_seqlens) == n_levels):
        raise ValueError("one entry per shared level is needed in every list")
    if n_levels == 0:
        return [], []
    backend = _prefix_backend()
    use_tc = q.dtype in _TC_DTYPES and d in _TC_HEAD_DIMS and backend != "rowwise"
    if n_levels == 1 or not use_tc or os.environ.get("HYDRAGEN_B200_PREFIX_GROUPED", "1") == "0":
        outs, lses = [], []
        per_level = max(1, max_partials // n_levels)
        for i in range(n_levels):
            o, l = prefix_attention_partials(q, shared_ks[i], shared_vs[i], int(n_groups[i]), cu_seqlens[i], max_seqlens[i], max_splits=per_level)
            outs += o
            lses += l
        return outs, lses
    sm_scale = d**-0.5  # hydragen/flash.py:293
    hkv = shared_ks[0].shape[-2]
    outs = [torch.empty((b, nq, hq, d), device=q.device, dtype=q.dtype) for _ in range(n_levels)]
    lses = [torch.empty((b, nq, hq), device=q.device, dtype=torch.float32) for _ in range(n_levels)]
    q_rs = _rows_view(q)
    if q_rs is None:
        q = q.contiguous()
        q_rs = hq * d
    descs, keep = [], []
    for i in range(n_levels):
        k, v, ng, cu, mx = shared_ks[i], shared_vs[i], int(n_groups[i]), cu_seqlens[i], max_seqlens[i]
        _check_level(q, k, v, ng, cu, i)
        if k.shape[-2] != hkv:
            raise ValueError("all shared levels must have the same number of kv heads")
        k, v, n_k_rows, k_len, kv_rs, max_k = _level_views(q, k, v, ng, cu, mx, hkv, d)
        keep += [k, v]
        descs.append(_lib.make_prefix_level(k, v, outs[i], lses[i], cu, n_k_rows, kv_rs, ng, k_len, max_k if cu is not None else 0))
    # one launch covers up to MAX_PREFIX_LEVELS levels (deeper hierarchies: one launch per chunk of levels)
    for i in range(0, len(descs), _lib.MAX_PREFIX_LEVELS):
        _lib.prefix_attn_grouped_fwd(q, b * nq, q_rs, descs[i : i + _lib.MAX_PREFIX_LEVELS], hq, hkv, d, sm_scale, split=_prefix_split())
    return outs, lses


def causal_attention_tc(q: Tensor, k: Tensor, v: Tensor) -> Optional[Tuple[Tensor, Tensor]]:
    """Causal self-attention of a prefill chunk on the tcgen05 kernel when the shape allows it (16-bit, d in {64, 128},
    sk >= sq, at least one key block of queries): ``(out [b, sq, hq, d], lse [b, sq, hq])``; None otherwise (the
    caller then uses the CUDA-core kernel: fp32, other head dims, chunks too short to be worth a tensor-core launch)."""
    b, sq, hq, d = q.shape
    sk = k.shape[1]
    backend = _causal_backend()
    use_tc = q.dtype in _TC_DTYPES and d in _TC_HEAD_DIMS and sk >= sq and backend != "rowwise" and (backend == "tcgen05" or sq >= 64)
    if backend == "tcgen05" and not use_tc:
        raise ValueError(f"tcgen05 causal kernel does not take dtype {q.dtype} / head_dim {d} / sq {sq} > sk {sk}")
    if not use_tc:
        return None
    q_rs = _rows_view(q)
    if q_rs is None:
        q = q.contiguous()
        q_rs = hq * d
    kv_rs = _rows_view(k)
    if kv_rs is None or _rows_view(v) != kv_rs:
        k, v = k.contiguous(), v.contiguous()
        kv_rs = k.shape[2] * d
    out = torch.empty((b, sq, hq, d), device=q.device, dtype=q.dtype)
    lse = torch.empty((b, sq, hq), device=q.device, dtype=torch.float32)
    _lib.causal_attn_fwd(q, k, v, out, lse, b, sq, sk, hq, k.shape[2], d, q_rs, kv_rs, d**-0.5)
    return out, lse


def flash_attention(q: Tensor, k: Tensor, v: Tensor, causal: bool = False) -> Tuple[Tensor, Tensor]:
    """hydragen/flash.py:284-306.  q [b, sq, hq, d]; k, v [b, sk, hkv, d] -> out [b, sq, hq, d],
    lse [b, hq, sq] (fp32).  ``causal`` is bottom-right aligned when sq != sk (flash-attn >= 2.1)."""
    _check_qkv(q, k, v)
    b, sq, hq, d = q.shape
    if not causal:
        out, lse = prefix_attention_grouped(q, k, v, n_groups=b)
        return out, lse.permute(0, 2, 1)
    # prefill chunks of at least one key block go to the tensor cores; shorter ones (launch-bound either way; the
    # CUDA-core kernel keeps P in fp32) and shapes the tcgen05 kernel does not take stay on the CUDA-core kernel
    res = causal_attention_tc(q, k, v)
    if res is None:
        res = _rowwise(q, k, v, None, causal=True)
    return res[0], res[1].permute(0, 2, 1)


def _varlen_by_sequence(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q: int, causal: bool):
    """flash_attention_varlen for ragged query groups / causal masks: sequence i = q rows [cu_q[i], cu_q[i+1]) against key rows
    [cu_k[i], cu_k[i+1]), one ``flash_attention`` call each; lse [n, hq, max_seqlen_q] with -inf past a sequence's rows."""
    cq, ck = [int(x) for x in cu_seqlens_q.tolist()], [int(x) for x in cu_seqlens_k.tolist()]
    n = len(cq) - 1
    if len(ck) != n + 1:
        raise ValueError("cu_seqlens_q and cu_seqlens_k must describe the same number of sequences")
    total_q, hq, d = q.shape
    out = torch.zeros((total_q, hq, d), device=q.device, dtype=q.dtype)
    lse = torch.full((n, hq, max_seqlen_q), float("-inf"), device=q.device, dtype=torch.float32)
    for i in range(n):
        qs, qe, ks, ke = cq[i], cq[i + 1], ck[i], ck[i + 1]
        if qe - qs > max_seqlen_q:
            raise ValueError(f"sequence {i} has {qe - qs} query rows, max_seqlen_q is {max_seqlen_q}")
        if qe == qs or ke == ks:  # no rows / no keys: out = 0, lse = -inf
            continue
        o, l = flash_attention(q[qs:qe].unsqueeze(0), k[ks:ke].unsqueeze(0), v[ks:ke].unsqueeze(0), causal=causal)
        out[qs:qe] = o[0]
        lse[i, :, : qe - qs] = l[0]
    return out, lse


def flash_attention_varlen(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    cu_seqlens_q: Tensor,
    cu_seqlens_k: Tensor,
    max_seqlen_q: int,
    max_seqlen_k: int,
    causal: bool = False,
) -> Tuple[Tensor, Tensor]:
    """hydragen/flash.py:309-351.  q [total_q, hq, d], k/v [total_k, hkv, d]; returns
    out [total_q, hq, d] and lse [n, hq, max_seqlen_q] (the flash-attn v2.3.6 layout).

    Every call the reference makes (hydragen/attention.py:295-321) has n equal query groups of ``max_seqlen_q`` rows and no
    mask: that form is ONE grouped tcgen05 launch, recognised without a device sync (n * max_seqlen_q == total_q), the
    ragged key side read from ``cu_seqlens_k`` on the device.  The general form of the primitive -- ragged query groups and / or
    ``causal=True`` (bottom-right aligned per sequence, flash-attn >= 2.1) -- is kept for surface parity and runs one dense call
    per sequence after reading the offsets on the host (a sync; not a hot path: nothing on the Hydragen path produces it)."""
    if q.ndim != 3 or k.ndim != 3 or v.ndim != 3:
        raise ValueError("varlen tensors must be [total, heads, d]")
    n = cu_seqlens_q.shape[0] - 1
    total_q, hq, d = q.shape
    if causal or n * max_seqlen_q != total_q:
        return _varlen_by_sequence(q, k, v, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, causal)
    out, lse = prefix_attention_grouped(q.view(n, max_seqlen_q, hq, d) if q.is_contiguous() else q.reshape(n, max_seqlen_q, hq, d),
                                        k, v, n_groups=n, cu_seqlens_k=cu_seqlens_k, max_seqlen_k=max_seqlen_k)
    return out.view(total_q, hq, d), lse.view(n, max_seqlen_q, hq).permute(0, 2, 1)


def _rowwise(q, k, v, seq_len, causal, partial_outs=(), partial_lses=()):
    b, nq, hq, d = q.shape
    if not _inner_ok(q):
        q = q.contiguous()
    if not _inner_ok(k):
        k = k.contiguous()
    if not _inner_ok(v) or v.stride() != k.stride():
        v = v.contiguous()
        k = k.contiguous()
    out = torch.empty((b, nq, hq, d), device=q.device, dtype=q.dtype)
    lse = torch.empty((b, nq, hq), device=q.device, dtype=torch.float32)
    _lib.rowwise_attn_fwd(q, k, v, seq_len, None, 1, causal, out, lse, k.shape[1],
                          (k.stride(0), k.stride(1), k.stride(2)), list(partial_outs), list(partial_lses), d**-0.5)
    return out, lse


def flash_attention_seqlen(raw_q: Tensor, raw_k: Tensor, raw_v: Tensor, seq_len: Optional[Tensor] = None):
    """hydragen/flash.py:163-281.  q [b, q, hq, d]; k, v [b, kmax, hkv, d]; sequence b attends to
    keys < seq_len[b] (int32 or int64, on device; no causal mask).  Returns out [b, q, hq, d] and
    lse [b, q, hq] fp32.  (``seq_len=None`` crashes in the reference; here it means "all keys".)"""
    _check_qkv(raw_q, raw_k, raw_v)
    return _rowwise(raw_q, raw_k, raw_v, seq_len, causal=False)


def suffix_attention_fused(q, k, v, seq_len, causal, partial_outs, partial_lses):
    """Suffix branch + combine in one launch (hydragen/attention.py:343-352): the per-sequence
    result is merged in registers with the prefix partials; returns (final out, merged lse)."""
    return _rowwise(q, k, v, seq_len, causal, partial_outs, partial_lses)


def decode_attention_fused(q: Tensor, k_new: Tensor, v_new: Tensor, positions: Tensor, k_cache: Tensor, v_cache: Tensor,
                           partial_outs=(), partial_lses=()) -> Tuple[Tensor, Tensor]:
    """One decode step of the suffix side in one launch: append ``k_new / v_new`` ``[b, 1, hkv, d]`` at row
    ``positions[b]`` of the unique caches ``[b_max, lk, hkv, d]``, attend ``q [b, 1, hq, d]`` to the
    ``positions[b] + 1`` keys of its own sequence and merge with the prefix partials.  Replaces
    ``update_per_completion_kvs`` + ``flash_attention_seqlen`` + ``combine_lse`` of the reference's decode
    branch (hydragen/llama.py:565-587).  Returns (out, merged lse [b, 1, hq])."""
    _check_qkv(q, k_new, v_new)
    b, nq, hq, d = q.shape
    if nq != 1 or k_new.shape[1] != 1:
        raise ValueError("decode_attention_fused takes exactly one new token per sequence")
    if k_cache.shape != v_cache.shape or k_cache.stride() != v_cache.stride() or k_cache.stride(-1) != 1:
        raise ValueError("k_cache / v_cache must share shape and strides, head_dim contiguous")
    if k_cache.shape[0] < b or k_cache.shape[2:] != k_new.shape[2:] or k_cache.dtype != q.dtype:
        raise ValueError(f"cache {tuple(k_cache.shape)} does not match new rows {tuple(k_new.shape)}")
    if not _inner_ok(q):
        q = q.contiguous()
    positions = positions.reshape(-1)
    if positions.shape[0] != b:
        raise ValueError("positions must hold one row index per sequence")
    out = torch.empty((b, 1, hq, d), device=q.device, dtype=q.dtype)
    lse = torch.empty((b, 1, hq), device=q.device, dtype=torch.float32)
    _lib.decode_attn_fused(q, k_new.contiguous(), v_new.contiguous(), positions.contiguous(), k_cache, v_cache, out, lse,
                           list(partial_outs), list(partial_lses), d**-0.5)
    return out, lse
