"""Test helper: swap the attention entry points that ``hydragen_b200.llama`` calls for the CPU oracle
(fp64), so that the SAME model code can be run once on the CUDA kernels and once on the checker.  This is
how the e2e parity tests stand in for the reference's HuggingFace comparison (tests/test_e2e.py:91-119 of the
reference), which needs hub weights that do not exist offline."""

import torch

from oracle import hydragen_oracle as O


def _back(x, like):
    return x.to(device=like.device, dtype=like.dtype)


def _c(x):
    if x is None:
        return None
    if isinstance(x, (list, tuple)):
        return [_c(t) for t in x]
    return x.detach().cpu() if isinstance(x, torch.Tensor) else x


def flash_attention(q, k, v, causal=False):
    out, lse = O.flash_attention(_c(q), _c(k), _c(v), causal=causal)
    return _back(out, q), lse.to(q.device, torch.float32)


def flash_attention_seqlen(q, k, v, seq_len=None):
    out, lse = O.flash_attention_seqlen(_c(q), _c(k), _c(v), seq_len=_c(seq_len))
    return _back(out, q), lse.to(q.device, torch.float32)


def hydragen_attention(q, k, v, shared_ks, shared_vs, shared_cu_seq_lens, shared_max_seq_lens, use_varlens, seq_lens=None):
    out = O.hydragen_attention(_c(q), _c(k), _c(v), _c(shared_ks), _c(shared_vs), _c(shared_cu_seq_lens), shared_max_seq_lens, use_varlens,
                               seq_lens=_c(seq_lens))
    return _back(out, q)


def kv_append(k_new, v_new, positions, k_cache, v_cache):
    """hydragen/llama.py:250-257: scatter_ with a fully expanded index."""
    b, s, h, d = k_new.shape
    idx = positions.view(b, s, 1, 1).expand(b, s, h, d).long()
    k_cache[:b].scatter_(1, idx, k_new)
    v_cache[:b].scatter_(1, idx, v_new)


def hydragen_attention_decode(q, k_new, v_new, positions, k_cache, v_cache, shared_ks, shared_vs, shared_cu_seq_lens=None,
                              shared_max_seq_lens=None, use_varlens=None):
    """The reference's decode branch (hydragen/llama.py:564-587): append, then attention with seq_len = pos + 1."""
    n = len(shared_ks)
    kv_append(k_new, v_new, positions, k_cache, v_cache)
    b = q.shape[0]
    seq_lens = positions.reshape(-1)[:b] + 1
    if n == 0:
        return flash_attention_seqlen(q, k_cache[:b], v_cache[:b], seq_lens)[0]
    return hydragen_attention(q, k_cache[:b], v_cache[:b], shared_ks, shared_vs, shared_cu_seq_lens or [None] * n,
                              shared_max_seq_lens or [None] * n, use_varlens or [False] * n, seq_lens=seq_lens)


def apply_rotary_pos_emb(q, k, cos, sin, position_ids, unsqueeze_dim=2, inplace=False):
    """hydragen/llama.py:494-501 through the oracle's eager restatement (device-agnostic torch ops)."""
    return O.apply_rotary_pos_emb(q, k, cos, sin, position_ids, unsqueeze_dim=unsqueeze_dim)


def apply(monkeypatch):
    """Route every attention call of hydragen_b200.llama through the oracle."""
    import hydragen_b200.llama as L

    monkeypatch.setattr(L, "flash_attention", flash_attention)
    monkeypatch.setattr(L, "flash_attention_seqlen", flash_attention_seqlen)
    monkeypatch.setattr(L, "hydragen_attention", hydragen_attention)
    monkeypatch.setattr(L, "hydragen_attention_decode", hydragen_attention_decode)
    monkeypatch.setattr(L, "kv_append", kv_append)
    monkeypatch.setattr(L, "apply_rotary_pos_emb", apply_rotary_pos_emb)
