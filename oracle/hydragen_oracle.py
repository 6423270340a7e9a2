"""CPU oracle for the Hydragen shared-prefix attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``hydragen_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker (or as
the timed CPU baseline), never as the product path.

What it restates (all citations are into /root/reference):

* ``attention_ref``            the softmax-attention primitive the reference obtains
                               from flash-attn v2.3.6 (requirements.txt:7, NOT in tree;
                               call sites hydragen/flash.py:295-304, 336-349).  Published
                               semantics relied on: scale d^-1/2 passed explicitly
                               (flash.py:293), LSE = natural log of sum exp(scaled scores) in
                               fp32 laid out [b, h, sq], GQA by head ratio
                               (q head h -> kv head h // (Hq/Hkv)), causal mask bottom-right
                               aligned when sq != sk (flash-attn >= 2.1).
* ``flash_attention`` / ``flash_attention_varlen`` / ``flash_attention_seqlen``
                               same return layouts as hydragen/flash.py:284-306, 309-351,
                               163-281 (the last one: keys < seq_len[b], no causal mask,
                               LSE [b, q, h]: flash.py:273-281, xformers_stuff.py:274-279).
* ``combine_lse_torch``        hydragen/attention.py:21-43.
* ``hydragen_attention``       hydragen/attention.py:177-354 (decomposition, inter-sequence
                               batching order "(n s)", LSE layout [b, nq, h], early return
                               when k is empty and there is one level).
* ``concat_attention``         ground truth of tests/test_attention.py:132-178.
* ``rotary_tables`` / ``apply_rotary_pos_emb``
                               the step right before the path: RoPE as
                               hydragen/llama.py:47-55, 494-501 calls it.  The arithmetic lives in
                               transformers==4.37.2 (requirements.txt, NOT in tree): half-split
                               rotation, tables cos/sin(cat(freqs, freqs)), inv_freq = base^(-2i/d),
                               ``cos[position_ids].unsqueeze(dim)``.  Pinned against the installed
                               transformers' own ``apply_rotary_pos_emb`` / ``rotate_half`` (same
                               formula, gather done by the caller since v4.38):
                               tests/golden/make_golden_rope.py -> tests/golden/rope_golden.npz.

Pinning: the reference keeps NO golden vectors for this path and its
attention entry points cannot execute without a GPU.  The oracle is pinned instead
against the reference ITSELF run in the build container: ``tests/golden/make_golden.py``
imports /root/reference/hydragen/attention.py, swaps only the three third-party
primitives (flash-attn / Triton, which need CUDA) for the ones below, and runs the
reference's own ``hydragen_attention`` / ``combine_lse_torch`` code on the reference's
own test case list; the committed outputs are what ``tests/test_oracle.py`` checks.
The isolated primitive ``(out, lse)`` is pinned by the mathematical definition only
(the reference's tests never pin flash-attn against an independent implementation).

Precision: ``compute_dtype`` is torch.float64 for goldens and parity checks and
torch.float32 for the timed CPU baseline.  ``round_to`` optionally reproduces the
reference's rounding points (branch outputs rounded to the q dtype before the
combine: README.md:488-490).
"""

from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
from torch import Tensor

__all__ = [
    "rdiff",
    "attention_ref",
    "flash_attention",
    "flash_attention_varlen",
    "flash_attention_seqlen",
    "combine_lse_torch",
    "hydragen_attention",
    "hydragen_attention_nopad",
    "rotary_tables",
    "rotate_half",
    "apply_rotary_pos_emb",
    "concat_attention",
]


def rdiff(a: Tensor, b: Tensor, eps: float = 1e-8) -> Tensor:
    """hydragen/utils.py:13-15 -- the parity metric of every reference test."""
    diff = (a - b).abs()
    return 2 * diff / (a.abs() + b.abs() + eps)


# ---------------------------------------------------------------------------
# primitive
# ---------------------------------------------------------------------------


def attention_ref(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    causal: bool = False,
    seq_len: Optional[Tensor] = None,
    compute_dtype: torch.dtype = torch.float64,
):
    """softmax(q k^T / sqrt(d)) v with its log-sum-exp.

    q [b, sq, hq, d]; k, v [b, sk, hkv, d].  Returns out [b, sq, hq, d] and
    lse [b, hq, sq], both in ``compute_dtype``.  A query row with no valid key gets
    out = 0 and lse = -inf (the convention of the CUDA kernels; the reference never
    reaches that case: SURVEY.md A.2).
    """
    b, sq, hq, d = q.shape
    bk, sk, hkv, dk = k.shape
    assert bk == b and dk == d and v.shape == k.shape, (q.shape, k.shape, v.shape)
    assert hq % hkv == 0
    g = hq // hkv
    scale = d**-0.5  # flash.py:293

    qf = q.to(compute_dtype).permute(0, 2, 1, 3)  # b hq sq d
    kf = k.to(compute_dtype).permute(0, 2, 1, 3)  # b hkv sk d
    vf = v.to(compute_dtype).permute(0, 2, 1, 3)
    # q head h = kh * g + j  ->  kv head kh   (flash.py:176 "(kh qh)")
    qf = qf.reshape(b, hkv, g * sq, d)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale  # b hkv (g sq) sk
    s = s.reshape(b, hkv, g, sq, sk)

    mask = None
    if seq_len is not None:
        sl = seq_len.to(torch.int64).reshape(b, 1, 1, 1, 1)
        mask = torch.arange(sk).reshape(1, 1, 1, 1, sk) < sl
    if causal:
        # bottom-right aligned: query i sees keys j <= i + (sk - sq)
        qi = torch.arange(sq).reshape(1, 1, 1, sq, 1)
        kj = torch.arange(sk).reshape(1, 1, 1, 1, sk)
        cm = kj <= qi + (sk - sq)
        mask = cm if mask is None else (mask & cm)
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))

    if sk == 0:
        lse = torch.full((b, hkv, g, sq), float("-inf"), dtype=compute_dtype)
        out = torch.zeros((b, hkv, g, sq, d), dtype=compute_dtype)
    else:
        m = s.max(dim=-1, keepdim=True).values
        m_safe = torch.where(torch.isinf(m), torch.zeros_like(m), m)
        p = torch.exp(s - m_safe)
        l = p.sum(dim=-1, keepdim=True)
        lse = (m_safe + torch.log(l)).squeeze(-1)  # log(0) = -inf for empty rows
        p = p / torch.where(l == 0, torch.ones_like(l), l)
        out = torch.matmul(p.reshape(b, hkv, g * sq, sk), vf).reshape(b, hkv, g, sq, d)

    out = out.reshape(b, hq, sq, d).permute(0, 2, 1, 3).contiguous()
    lse = lse.reshape(b, hq, sq).contiguous()
    return out, lse


def flash_attention(q, k, v, causal: bool = False, compute_dtype=torch.float64, round_to=None):
    """Layouts of hydragen/flash.py:284-306: out [b,sq,hq,d], lse fp32-like [b,hq,sq]."""
    out, lse = attention_ref(q, k, v, causal=causal, compute_dtype=compute_dtype)
    if round_to is not None:
        out = out.to(round_to)
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
    compute_dtype=torch.float64,
    round_to=None,
):
    """hydragen/flash.py:309-351.  q [total_q, hq, d], k/v [total_k, hkv, d];
    returns out [total_q, hq, d] and lse [n, hq, max_seqlen_q] (the flash-attn v2.3.6
    layout consumed at attention.py:333-338; rows past a group's q length are -inf... the
    reference never reads them because every group has exactly max_seqlen_q queries)."""
    n = cu_seqlens_q.shape[0] - 1
    tq, hq, d = q.shape
    out = torch.zeros((tq, hq, d), dtype=compute_dtype)
    lse = torch.full((n, hq, max_seqlen_q), float("-inf"), dtype=compute_dtype)
    cq = [int(x) for x in cu_seqlens_q]
    ck = [int(x) for x in cu_seqlens_k]
    for i in range(n):
        qs, qe, ks, ke = cq[i], cq[i + 1], ck[i], ck[i + 1]
        assert qe - qs <= max_seqlen_q and ke - ks <= max_seqlen_k
        o, l = attention_ref(
            q[qs:qe].unsqueeze(0), k[ks:ke].unsqueeze(0), v[ks:ke].unsqueeze(0),
            causal=causal, compute_dtype=compute_dtype,
        )
        out[qs:qe] = o[0]
        lse[i, :, : qe - qs] = l[0]
    if round_to is not None:
        out = out.to(round_to)
    return out, lse


def flash_attention_seqlen(q, k, v, seq_len=None, compute_dtype=torch.float64, round_to=None):
    """hydragen/flash.py:163-281: keys < seq_len[b] are valid, no causal mask;
    returns out [b, q, hq, d] and lse [b, q, hq] (flash.py:273-281)."""
    out, lse = attention_ref(q, k, v, causal=False, seq_len=seq_len, compute_dtype=compute_dtype)
    if round_to is not None:
        out = out.to(round_to)
    return out, lse.permute(0, 2, 1).contiguous()


# ---------------------------------------------------------------------------
# combine
# ---------------------------------------------------------------------------


def combine_lse_torch(outs: Sequence[Tensor], lses: Sequence[Tensor], out_dtype=None):
    """hydragen/attention.py:21-43.  outs n x [b,s,h,d], lses n x [b,s,h].
    Differs from the reference only in guarding an all-(-inf) row (-> 0) instead of NaN."""
    o = torch.stack(list(outs))
    l = torch.stack(list(lses)).to(o.dtype if o.dtype in (torch.float32, torch.float64) else torch.float32)
    m = l.max(0).values
    m = torch.where(torch.isinf(m) & (m < 0), torch.zeros_like(m), m)
    w = (l - m[None]).exp()
    den = w.sum(0)
    den = torch.where(den == 0, torch.ones_like(den), den)
    agg = (o * w.unsqueeze(-1)).sum(0) / den.unsqueeze(-1)
    return agg.to(out_dtype if out_dtype is not None else o.dtype)


# ---------------------------------------------------------------------------
# the operator
# ---------------------------------------------------------------------------


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
    compute_dtype=torch.float64,
    round_to=None,
):
    """Restatement of hydragen/attention.py:177-354 (same argument meaning)."""
    assert q.ndim == 4 and k.ndim == 4 and v.ndim == 4
    assert k.shape == v.shape
    assert len(shared_ks) == len(shared_vs) == len(shared_cu_seq_lens) == len(shared_max_seq_lens) == len(use_varlens)
    b, nq, hq, d = q.shape
    outs, lses = [], []
    for sk, sv, scu, smax, use_varlen in zip(shared_ks, shared_vs, shared_cu_seq_lens, shared_max_seq_lens, use_varlens):
        assert sk.shape == sv.shape
        if not use_varlen:
            n = sk.shape[0]
            assert b % n == 0
            s = b // n
            # "(n s) nq hq d -> n (s nq) hq d"  attention.py:264-268 (a view)
            bq = q.reshape(n, s * nq, hq, d)
            so, sl = flash_attention(bq, sk, sv, compute_dtype=compute_dtype, round_to=round_to)
            so = so.reshape(b, nq, hq, d)
            if k.shape[1] == 0 and len(shared_ks) == 1:
                return so  # attention.py:273-274
            # "n h (s nq) -> (n s) nq h"  attention.py:276-280
            sl = sl.reshape(n, hq, s, nq).permute(0, 2, 3, 1).reshape(b, nq, hq).contiguous()
        else:
            n = scu.shape[0] - 1
            assert b % n == 0
            s = b // n
            qps = s * nq
            bq = q.reshape(b * nq, hq, d)
            cu_q = torch.arange(0, n + 1, dtype=torch.int32) * qps  # attention.py:295-311
            so, sl = flash_attention_varlen(
                bq, sk, sv, cu_seqlens_q=cu_q, cu_seqlens_k=scu,
                max_seqlen_q=qps, max_seqlen_k=smax, compute_dtype=compute_dtype, round_to=round_to,
            )
            so = so.reshape(b, nq, hq, d)
            if k.shape[1] == 0 and len(shared_ks) == 1:
                return so  # attention.py:330-331
            sl = sl.reshape(n, hq, s, nq).permute(0, 2, 3, 1).reshape(b, nq, hq).contiguous()
        outs.append(so)
        lses.append(sl)

    if seq_lens is None:
        uo, ul = flash_attention(q, k, v, causal=True, compute_dtype=compute_dtype, round_to=round_to)
        ul = ul.permute(0, 2, 1).contiguous()  # "b h q -> b q h"  attention.py:345
    else:
        uo, ul = flash_attention_seqlen(q, k, v, seq_len=seq_lens, compute_dtype=compute_dtype, round_to=round_to)
    outs.append(uo)
    lses.append(ul)
    outs = [o.to(compute_dtype) for o in outs]
    return combine_lse_torch(outs, lses, out_dtype=round_to)


def hydragen_attention_nopad(q, k, v, shared_ks, shared_vs, seq_len=None, **kw):
    """hydragen/attention.py:357-392."""
    n = len(shared_ks)
    return hydragen_attention(q, k, v, shared_ks, shared_vs, [None] * n, [None] * n, [False] * n, seq_lens=seq_len, **kw)


def concat_attention(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    shared_ks: List[Tensor],
    shared_vs: List[Tensor],
    shared_cu_seq_lens: List[Optional[Tensor]],
    use_varlens: List[bool],
    seq_lens: Optional[Tensor] = None,
    compute_dtype=torch.float64,
):
    """Ground truth of tests/test_attention.py:132-178: for each sequence, attention over the
    explicit concatenation [shared level 0 rows of its parent, level 1 ..., own k[:len]],
    shared parent index = i // (B / sb)."""
    b = q.shape[0]
    res = []
    for i in range(b):
        ks, vs = [], []
        for sk, sv, scu, uv in zip(shared_ks, shared_vs, shared_cu_seq_lens, use_varlens):
            if uv:
                sb = scu.shape[0] - 1
                idx = i // (b // sb)
                s0, s1 = int(scu[idx]), int(scu[idx + 1])
                ks.append(sk[s0:s1].unsqueeze(0))
                vs.append(sv[s0:s1].unsqueeze(0))
            else:
                sb = sk.shape[0]
                idx = i // (b // sb)
                ks.append(sk[idx].unsqueeze(0))
                vs.append(sv[idx].unsqueeze(0))
        uk, uv_ = k[i : i + 1], v[i : i + 1]
        if seq_lens is not None:
            uk, uv_ = uk[:, : int(seq_lens[i])], uv_[:, : int(seq_lens[i])]
        kk = torch.cat(ks + [uk], dim=1)
        vv = torch.cat(vs + [uv_], dim=1)
        # nq == 1 in the reference test, so non-causal == causal here; for nq > 1 the
        # decomposition's suffix branch is causal (attention.py:344) and so is the truth.
        causal = seq_lens is None and q.shape[1] > 1
        o, _ = attention_ref(q[i : i + 1], kk, vv, causal=causal, compute_dtype=compute_dtype)
        res.append(o)
    return torch.cat(res, dim=0)


# ---------------------------------------------------------------------------
# case builders shared by the golden generator, the tests and the bench
# ---------------------------------------------------------------------------

# tests/test_attention.py:26-32, verbatim as data.
REFERENCE_SIZES_LIST = [
    [[1], [10]],
    [[3], [6, 6]],
    [[3], [6, 7]],
    [[7, 7], [9, 10, 11, 4], [129, 2, 3, 4, 5, 6, 7, 128]],
    [[16384], [1, 128, 256]],
]


def build_case(sizes, qheads: int, kvheads: int, dim: int, dtype=torch.float16, seed: int = 0, nq: int = 1):
    """Inputs of one tests/test_attention.py case (:40-112), seeded on the CPU generator
    so that the same tensors can be regenerated on any box with this image."""
    g = torch.Generator().manual_seed(seed)

    def randn(*shape):
        return torch.randn(*shape, generator=g, dtype=torch.float32).to(dtype)

    bsz = len(sizes[-1])
    q = randn(bsz, nq, qheads, dim)
    shared_ks, shared_vs, shared_cu, max_lens, use_varlens = [], [], [], [], []
    for lens in sizes[:-1]:
        uv = len(set(lens)) > 1
        use_varlens.append(uv)
        if uv:
            sk = randn(sum(lens), kvheads, dim)
            cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
            cu[1:] = torch.tensor(lens, dtype=torch.int32).cumsum(0)
            shared_cu.append(cu)
            max_lens.append(max(lens))
        else:
            sk = randn(len(lens), lens[0], kvheads, dim)
            shared_cu.append(None)
            max_lens.append(None)
        sv = randn(*sk.shape)
        shared_ks.append(sk)
        shared_vs.append(sv)
    final = sizes[-1]
    k = randn(bsz, max(final), kvheads, dim)
    v = randn(bsz, max(final), kvheads, dim)
    seq_lens = torch.tensor(final, dtype=torch.int32) if len(set(final)) > 1 else None
    return dict(
        q=q, k=k, v=v, shared_ks=shared_ks, shared_vs=shared_vs, shared_cu_seq_lens=shared_cu,
        shared_max_seq_lens=max_lens, use_varlens=use_varlens, seq_lens=seq_lens,
    )


# ----------------------------------------------------------------------------------------------
# RoPE (the step before the path)
# ----------------------------------------------------------------------------------------------


def rotary_tables(dim: int, max_position_embeddings: int, base: float = 10000.0, dtype=torch.float32):
    """LlamaRotaryEmbedding of transformers 4.37.2 as imported at hydragen/llama.py:1-10 (pinned upstream, not in
    tree): inv_freq = base^(-2i/d), emb = cat(freqs, freqs); HydragenLlamaRotaryEmbedding.forward
    (hydragen/llama.py:47-55) returns the full tables cast to the activation dtype."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    t = torch.arange(max_position_embeddings, dtype=torch.float32)
    emb = torch.cat((torch.outer(t, inv_freq),) * 2, dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x: Tensor) -> Tensor:
    half = x.shape[-1] // 2
    return torch.cat((-x[..., half:], x[..., :half]), dim=-1)


def apply_rotary_pos_emb(q: Tensor, k: Tensor, cos: Tensor, sin: Tensor, position_ids: Tensor, unsqueeze_dim: int = 2):
    """The call at hydragen/llama.py:494-501: q [b, s, hq, d], k [b, s, hkv, d], cos/sin [max_pos, d] in the
    activation dtype, position_ids [b, s] absolute positions.  Evaluated eagerly in the activation dtype,
    operation by operation, exactly as transformers 4.37.2 does (the CUDA kernel reproduces the same bits)."""
    c = cos[position_ids].unsqueeze(unsqueeze_dim)
    s_ = sin[position_ids].unsqueeze(unsqueeze_dim)
    return (q * c) + (rotate_half(q) * s_), (k * c) + (rotate_half(k) * s_)
