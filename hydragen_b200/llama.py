"""Llama decode path around the Hydragen attention operator -- the caller boundary of the hot path.

Same public surface as the reference's ``hydragen/llama.py`` (class / method / argument names and
their meaning): ``SharedCache``, ``PerLayerKVCache``, ``AttentionMode``,
``hydragen_attention_on_caches``, ``HydragenLlamaAttention``, ``HydragenLlamaForCausalLM`` with
``setup_caches / graph / append_shared / process_unique / generate / empty_shared_cache /
truncate_shared_caches``, ``SharedCacheOp``.  State-dict keys are the HuggingFace Llama ones, so
``load_state_dict(hf_model.state_dict())`` works (hydragen/llama.py:1398-1422).

What is different from the reference (B200-first, not a port):

* attention goes through the sm_100a kernels behind the C ABI (``hydragen_b200.attention`` /
  ``.flash``); per decode step and layer that is 1 KV-append launch + L prefix launches + 1 fused
  suffix/combine launch (reference: 2 ``scatter_`` with a materialised int64 index, flash-attn,
  LSE transposes, int cast, split-K, reduce, combine -- hydragen/llama.py:236-262, 564-587);
* everything that depends only on ``position_ids`` (RoPE cos/sin rows, the per-sequence shared
  length, the unique position, the decode ``seq_lens``) is computed ONCE per forward and handed to
  the layers, instead of once per layer (hydragen/llama.py:485-501, 317-330, 569);
* the unique KV cache of all layers is one allocation ``[layers, 2, B, Lu, Hkv, d]`` sized for the
  180 GB HBM of a B200; layers hold views;
* transformers / accelerate are not imported: the Llama building blocks (RMSNorm, SwiGLU MLP,
  half-split RoPE -- "pinned upstream, not in tree": transformers==4.37.2 modeling_llama) are
  restated in plain PyTorch.  These, the linear layers (cuBLAS) and sampling are plumbing around
  the hot path and are deliberately not hand-written kernels.

There is no CPU fallback: the attention entry points raise on non-CUDA tensors.  Tests swap the
module-level names ``flash_attention / flash_attention_seqlen / hydragen_attention / kv_append``
for CPU checkers to exercise the host logic without a GPU.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Union

import torch
from torch import Tensor, nn

from . import _lib
from .attention import hydragen_attention, hydragen_attention_decode
from .flash import flash_attention, flash_attention_seqlen
from .rope import apply_rotary_pos_emb

kv_append = _lib.kv_append  # (k_new, v_new, positions, k_cache, v_cache): module-level so tests can swap it


# ----------------------------------------------------------------------------------------------
# config + building blocks (plumbing)
# ----------------------------------------------------------------------------------------------


@dataclass
class LlamaConfig:
    """The fields of transformers' LlamaConfig that this path reads; any object with these
    attributes (including a HuggingFace LlamaConfig) is accepted wherever a config is taken."""

    vocab_size: int = 32000
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: Optional[int] = None
    max_position_embeddings: int = 4096
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    attention_bias: bool = False
    pad_token_id: Optional[int] = None
    rope_scaling: Optional[dict] = None
    head_dim: Optional[int] = None

    def __post_init__(self):
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads
        if self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_attention_heads


def llama_config(name: str, **overrides) -> LlamaConfig:
    """Named architectures used by BASELINE.json's configs (random-init; there are no weights offline)."""
    table = {
        "llama-2-7b": dict(hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32),
        "llama-2-13b": dict(hidden_size=5120, intermediate_size=13824, num_hidden_layers=40, num_attention_heads=40),
        "tiny": dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                     num_key_value_heads=2, vocab_size=512, max_position_embeddings=512),
    }
    kw = dict(table[name.lower()])
    kw.update(overrides)
    return LlamaConfig(**kw)


def _cfg_head_dim(config) -> int:
    hd = getattr(config, "head_dim", None)
    return int(hd) if hd else config.hidden_size // config.num_attention_heads


class LlamaRMSNorm(nn.Module):
    def __init__(self, hidden_size: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x: Tensor) -> Tensor:
        dt = x.dtype
        xf = x.float()
        xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.variance_epsilon)
        return self.weight * xf.to(dt)


def _rowwise_then_all_reduce(x: Tensor, lin: nn.Linear, all_reduce) -> Tensor:
    """Row-parallel projection + the sum of its partials over the tensor-parallel group (hydragen/tp.py:99,108-112): one
    fused launch where the collective offers it (tp._AllReduce.linear), else the two steps."""
    fused = getattr(all_reduce, "linear", None)
    return fused(x, lin) if fused is not None else all_reduce(lin(x))


class LlamaMLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.gate_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.up_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.down_proj = nn.Linear(config.intermediate_size, config.hidden_size, bias=False)
        self.all_reduce = None  # set by tp.apply_tp

    def forward(self, x: Tensor) -> Tensor:
        h = nn.functional.silu(self.gate_proj(x)) * self.up_proj(x)
        if self.all_reduce is not None:
            return _rowwise_then_all_reduce(h, self.down_proj, self.all_reduce)
        return self.down_proj(h)


class HydragenLlamaRotaryEmbedding(nn.Module):
    """cos/sin tables for half-split RoPE (emb = cat(freqs, freqs)), hydragen/llama.py:47-55.

    The tables are NOT module buffers: they are computed on first use, on the device they are asked for, in fp32
    (exactly as the reference's buffers are: on CPU-independent arithmetic) and cached per (dtype, device).  A model
    whose parameters were created on the ``meta`` device (``from_pretrained`` / ``from_pretrained_tp`` build the module
    tree there before the checkpoint is assigned) therefore never owns a meta buffer that ``.to(device)`` would have
    to copy -- the reference gets the same effect from ``init_empty_weights(include_buffers=False)``."""

    def __init__(self, dim: int, max_position_embeddings: int = 2048, base: float = 10000.0, scaling_factor: float = 1.0):
        super().__init__()
        self.dim, self.max_position_embeddings, self.base, self.scaling_factor = dim, max_position_embeddings, float(base), float(scaling_factor)
        self._fp32: dict = {}  # device -> (cos, sin) fp32 [max_pos, dim]
        self._cast: dict = {}  # (dtype, device) -> tables cast once

    def _tables_fp32(self, device: torch.device):
        hit = self._fp32.get(device)
        if hit is None:
            # built on the CPU in fp32 and copied: identical bits on every device and rank
            inv_freq = 1.0 / (self.base ** (torch.arange(0, self.dim, 2, dtype=torch.float32) / self.dim))
            t = torch.arange(self.max_position_embeddings, dtype=torch.float32) / self.scaling_factor
            freqs = torch.outer(t, inv_freq)
            emb = torch.cat((freqs, freqs), dim=-1)
            hit = (emb.cos().to(device), emb.sin().to(device))
            self._fp32 = {device: hit}
        return hit

    @property
    def cos_cached(self) -> Tensor:
        return self._tables_fp32(torch.device("cpu"))[0]

    @property
    def sin_cached(self) -> Tensor:
        return self._tables_fp32(torch.device("cpu"))[1]

    def forward(self, x: Tensor, seq_len=None):
        """The full tables in the activation dtype on ``x``'s device.  The reference casts them on every call (once
        per layer per step); here the cast copy is made once per (dtype, device) and kept."""
        return self.tables(x.dtype, x.device)

    def tables(self, dtype: torch.dtype, device: Union[str, torch.device, None] = None):
        device = torch.device(device) if device is not None else torch.device("cpu")
        key = (dtype, device)
        hit = self._cast.get(key)
        if hit is None:
            cos, sin = self._tables_fp32(device)
            hit = (cos.to(dtype=dtype).contiguous(), sin.to(dtype=dtype).contiguous())
            self._cast = {key: hit}
        return hit


def repeat_to_batch_size(tensors: Sequence[Tensor], target_batch_size: Optional[int] = None) -> List[Tensor]:
    """hydragen/llama.py:32-44: level tensors of batch sb are repeat-interleaved to the full batch
    (the batch is grouped contiguously by shared parent)."""
    if target_batch_size is None:
        target_batch_size = max(t.shape[0] for t in tensors)
    out = []
    for t in tensors:
        assert target_batch_size % t.shape[0] == 0, f"{target_batch_size} {t.shape[0]}"
        out.append(t.repeat_interleave(target_batch_size // t.shape[0], dim=0))
    return out


# ----------------------------------------------------------------------------------------------
# KV caches
# ----------------------------------------------------------------------------------------------


class SharedCache(nn.Module):
    """One level of shared prefixes, packed: K/V ``[max_batch * max_len, Hkv, d]`` holding the valid
    rows of every shared sequence back to back, ``seq_lens`` int32 ``[max_batch]`` and
    ``cumsum_lengths`` int32 ``[max_batch + 1]`` (hydragen/llama.py:58-170)."""

    k_cache: Tensor
    v_cache: Tensor
    seq_lens: Tensor
    cumsum_lengths: Tensor

    def __init__(self, max_batch_size: int, max_seq_length: int, num_heads: int, head_dim: int, dtype: torch.dtype,
                 device: torch.device):
        super().__init__()
        rows = max_batch_size * max_seq_length
        self.register_buffer("k_cache", torch.zeros((rows, num_heads, head_dim), dtype=dtype, device=device), persistent=False)
        self.register_buffer("v_cache", torch.zeros((rows, num_heads, head_dim), dtype=dtype, device=device), persistent=False)
        self.register_buffer("seq_lens", torch.zeros((max_batch_size,), dtype=torch.int32, device=device), persistent=False)
        self.register_buffer("cumsum_lengths", torch.zeros((max_batch_size + 1,), dtype=torch.int32, device=device), persistent=False)
        self.max_batch_size = max_batch_size
        self.max_sequence_length = max_seq_length
        self.current_batch_size = 0
        # the tcgen05 prefix kernel is the same code for uniform and ragged levels (it reads
        # cumsum_lengths on the device), so unlike hydragen/llama.py:104-115 nothing is lost by
        # varlen; the flag is kept because callers and the graph-invalidation logic read it.
        self.use_varlen = False
        self.sliced_sequence_length: Optional[int] = None
        self.max_used_length = 0  # host copy of max(seq_lens) of the current fill (capacity checks without a device read)

    def get_current_batch_size(self) -> int:
        return self.current_batch_size

    def fill(self, key_states: Tensor, value_states: Tensor, seq_lens: Tensor, host_lens: Optional[List[int]] = None):
        """key/value_states [bs, L, Hkv, d] (right padded), seq_lens [bs] valid lengths.
        ``host_lens`` (the same lengths as Python ints) avoids the device->host read when the
        caller already knows them; otherwise there is ONE sync per fill (prefill only)."""
        bs, L = key_states.shape[0], key_states.shape[1]
        if bs > self.max_batch_size:
            raise ValueError(f"Batch size {bs} exceeds max batch size {self.max_batch_size}")
        if L > self.max_sequence_length:
            raise ValueError(f"Sequence length {L} exceeds max sequence length {self.max_sequence_length}")
        lens = [int(x) for x in (host_lens if host_lens is not None else seq_lens.tolist())]
        total = sum(lens)
        if min(lens) == L:  # no padding: the packed layout is the input itself
            self.k_cache[:total].copy_(key_states.reshape(total, *key_states.shape[2:]))
            self.v_cache[:total].copy_(value_states.reshape(total, *value_states.shape[2:]))
        else:
            keep = torch.arange(L, device=key_states.device)[None, :] < seq_lens.to(key_states.device)[:, None]
            self.k_cache[:total].copy_(key_states[keep])
            self.v_cache[:total].copy_(value_states[keep])
        self.seq_lens[:bs].copy_(seq_lens.to(torch.int32))
        self.cumsum_lengths[0] = 0
        self.cumsum_lengths[1 : bs + 1].copy_(seq_lens.cumsum(0).to(torch.int32))
        self.use_varlen = max(lens) != min(lens)
        self.sliced_sequence_length = None if self.use_varlen else lens[0]
        self.max_used_length = max(lens)
        self.current_batch_size = bs

    def get_used_cumsum_lengths(self) -> Tensor:
        return self.cumsum_lengths[: self.current_batch_size + 1]

    def get_used_seq_lens(self) -> Tensor:
        return self.seq_lens[: self.current_batch_size]


class PerLayerKVCache(nn.Module):
    """Per-sequence ("unique") K/V ``[max_batch, max_len, Hkv, d]`` plus the shared levels of one
    layer (hydragen/llama.py:173-346).  ``storage`` lets the model hand in views of one big
    allocation instead of allocating per layer."""

    per_completion_k_cache: Tensor
    per_completion_v_cache: Tensor

    def __init__(self, max_unique_batch_size: int, max_unique_seq_length: int, max_shared_batch_sizes: List[int],
                 max_shared_seq_lengths: List[int], n_kv_heads: int, head_dim: int, device: torch.device, dtype: torch.dtype,
                 storage: Optional[Tensor] = None):
        super().__init__()
        shape = (max_unique_batch_size, max_unique_seq_length, n_kv_heads, head_dim)
        if storage is None:
            storage = torch.zeros((2, *shape), dtype=dtype, device=device)
        assert tuple(storage.shape) == (2, *shape), f"{tuple(storage.shape)} {shape}"
        self.register_buffer("per_completion_k_cache", storage[0], persistent=False)
        self.register_buffer("per_completion_v_cache", storage[1], persistent=False)
        self.shared_caches = nn.ModuleList(
            [SharedCache(sb, sl, n_kv_heads, head_dim, dtype, device) for sb, sl in zip(max_shared_batch_sizes, max_shared_seq_lengths)]
        )
        self.num_used_shared_caches = 0

    # -- shared levels -----------------------------------------------------------------------
    def empty_shared_cache(self):
        self.truncate_shared_caches(0)

    def get_num_total_shared_caches(self) -> int:
        return len(self.shared_caches)

    def truncate_shared_caches(self, new_num_shared_caches: int):
        assert new_num_shared_caches <= self.get_num_total_shared_caches(), f"{new_num_shared_caches} {self.get_num_total_shared_caches()}"
        self.num_used_shared_caches = new_num_shared_caches

    def get_used_shared_caches(self) -> List[SharedCache]:
        return list(self.shared_caches)[: self.num_used_shared_caches]

    def has_shared(self) -> bool:
        return self.num_used_shared_caches > 0

    def get_shared_len(self, final_batch_size: int) -> Tensor:
        """Total shared length seen by each of the ``final_batch_size`` sequences (int64 [B])."""
        if self.num_used_shared_caches == 0:
            return torch.zeros((final_batch_size,), dtype=torch.long, device=self.per_completion_k_cache.device)
        lens = [c.get_used_seq_lens().to(torch.long) for c in self.get_used_shared_caches()]
        return sum(repeat_to_batch_size(lens, final_batch_size))

    def append_shared(self, key_states: Tensor, value_states: Tensor, seq_lens: Tensor, host_lens: Optional[List[int]] = None):
        if self.num_used_shared_caches >= self.get_num_total_shared_caches():
            raise ValueError(f"No more available shared caches: {self.num_used_shared_caches} {self.get_num_total_shared_caches()}")
        self.shared_caches[self.num_used_shared_caches].fill(key_states, value_states, seq_lens, host_lens)
        self.num_used_shared_caches += 1

    # -- unique cache ------------------------------------------------------------------------
    def update_per_completion_kvs(self, input_pos: Tensor, k_val: Tensor, v_val: Tensor):
        """Write k/v_val [bs, s, Hkv, d] at rows input_pos [bs, s] of the unique cache (one
        hg_kv_append launch; the reference: two scatter_ with an expanded int64 index,
        hydragen/llama.py:236-262).  Returns the first bs sequences of the whole cache."""
        assert input_pos.shape[1] == k_val.shape[1], f"{input_pos.shape} {k_val.shape}"
        bs = k_val.shape[0]
        kv_append(k_val.contiguous(), v_val.contiguous(), input_pos.contiguous(), self.per_completion_k_cache, self.per_completion_v_cache)
        return self.per_completion_k_cache[:bs], self.per_completion_v_cache[:bs]

    @torch.no_grad()
    def copy_shared_to_unique(self, total_num_sequences: int):
        """No-sharing baseline (hydragen/llama.py:264-298): the single shared level is replicated
        into every sequence's unique cache."""
        assert self.num_used_shared_caches == 1, "Cannot copy shared without exactly one active shared cache"
        sc: SharedCache = self.shared_caches[0]
        sb = sc.get_current_batch_size()
        assert total_num_sequences % sb == 0
        rep = total_num_sequences // sb
        cu = sc.get_used_cumsum_lengths().tolist()
        for i in range(sb):
            n = cu[i + 1] - cu[i]
            self.per_completion_k_cache[i * rep : (i + 1) * rep, :n] = sc.k_cache[cu[i] : cu[i + 1]].unsqueeze(0)
            self.per_completion_v_cache[i * rep : (i + 1) * rep, :n] = sc.v_cache[cu[i] : cu[i + 1]].unsqueeze(0)

    @torch.no_grad()
    def repeat_per_completion_cache_for_num_samples(self, current_size: int, num_samples: int):
        if num_samples == 1:
            return
        n = current_size * num_samples
        self.per_completion_k_cache[:n] = self.per_completion_k_cache[:current_size].repeat_interleave(num_samples, 0)
        self.per_completion_v_cache[:n] = self.per_completion_v_cache[:current_size].repeat_interleave(num_samples, 0)


class AttentionMode:
    SHARED_PREFILL = "shared-prefill"
    UNIQUE_PREFILL = "unique-prefill"
    DECODE = "decode"


def hydragen_attention_on_caches(q: Tensor, k: Tensor, v: Tensor, shared_caches: List[SharedCache], seq_len: Optional[Tensor] = None):
    """Adapts cache objects to the operator's argument lists (hydragen/llama.py:355-414): uniform
    levels are passed as [sb, L, Hkv, d] views of the packed buffer, ragged levels as the packed
    buffer + cumsum_lengths + max length."""
    keys, values, cu, mx, uv = _shared_cache_args(shared_caches)
    return hydragen_attention(q, k, v, shared_ks=keys, shared_vs=values, shared_cu_seq_lens=cu, shared_max_seq_lens=mx,
                              use_varlens=uv, seq_lens=seq_len)


def _shared_cache_args(shared_caches: List["SharedCache"]):
    keys, values, cu, mx, uv = [], [], [], [], []
    for sc in shared_caches:
        if sc.use_varlen:
            keys.append(sc.k_cache)
            values.append(sc.v_cache)
            cu.append(sc.get_used_cumsum_lengths())
            mx.append(sc.max_sequence_length)
        else:
            sb, L = sc.get_current_batch_size(), sc.sliced_sequence_length
            keys.append(sc.k_cache[: sb * L].view(sb, L, *sc.k_cache.shape[1:]))
            values.append(sc.v_cache[: sb * L].view(sb, L, *sc.v_cache.shape[1:]))
            cu.append(None)
            mx.append(None)
        uv.append(sc.use_varlen)
    return keys, values, cu, mx, uv


def hydragen_decode_on_caches(q: Tensor, k_new: Tensor, v_new: Tensor, positions: Tensor, cache: "PerLayerKVCache",
                              shared_caches: List["SharedCache"]):
    """The DECODE branch of the reference (hydragen/llama.py:564-587: update_per_completion_kvs, then
    hydragen_attention_on_caches / flash_attention_seqlen with seq_len = position + 1) as one prefix launch
    per shared level + ONE fused launch (append + suffix + combine)."""
    keys, values, cu, mx, uv = _shared_cache_args(shared_caches)
    bs = q.shape[0]
    return hydragen_attention_decode(q, k_new, v_new, positions, cache.per_completion_k_cache[:bs], cache.per_completion_v_cache[:bs],
                                     keys, values, cu, mx, uv)


# ----------------------------------------------------------------------------------------------
# model
# ----------------------------------------------------------------------------------------------


@dataclass
class StepContext:
    """Per-forward quantities that depend only on position_ids; computed once, used by every layer."""

    cos: Tensor  # the full RoPE tables [max_pos, d] in the activation dtype
    sin: Tensor
    position_ids: Tensor  # [b, s] absolute positions (rows of the tables)
    unique_position_ids: Tensor  # [b, s] position inside the unique cache
    seq_lens: Optional[Tensor]  # decode: unique_position + 1, [b] int64
    host_lens: Optional[List[int]] = None  # shared prefill: valid lengths as Python ints
    shared_lens_dev: Optional[Tensor] = None  # shared prefill: valid lengths [b]
    unique_max_len: Optional[int] = None  # unique prefill without hydragen: max position + 1


class HydragenLlamaAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = _cfg_head_dim(config)
        self.num_key_value_heads = config.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.disable_hydragen = False
        self.disable_attention = False  # throughput upper-bound ablation (hydragen/llama.py:433-437)
        self.fused_decode = True  # decode steps use the fused append + suffix + combine launch
        bias = bool(getattr(config, "attention_bias", False))
        self.q_proj = nn.Linear(self.hidden_size, self.num_heads * self.head_dim, bias=bias)
        self.k_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=bias)
        self.v_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=bias)
        self.o_proj = nn.Linear(self.num_heads * self.head_dim, self.hidden_size, bias=bias)
        self.kv_cache: Optional[PerLayerKVCache] = None
        self.mode: Optional[str] = None
        self.all_reduce = None  # set by tp.apply_tp: sums the row-parallel o_proj partials

    def forward(self, hidden_states: Tensor, ctx: StepContext) -> Tensor:
        b, s, _ = hidden_states.shape
        q = self.q_proj(hidden_states).view(b, s, self.num_heads, self.head_dim)
        k = self.k_proj(hidden_states).view(b, s, self.num_key_value_heads, self.head_dim)
        v = self.v_proj(hidden_states).view(b, s, self.num_key_value_heads, self.head_dim)
        # one launch (the reference: a gather + ten elementwise launches, hydragen/llama.py:494-501); in place on
        # the fresh projection outputs.  Absolute positions: shared K is stored already rotated.
        q, k = apply_rotary_pos_emb(q, k, ctx.cos, ctx.sin, ctx.position_ids, unsqueeze_dim=2, inplace=True)
        cache = self.kv_cache

        if self.disable_attention:
            out = q
        elif self.mode == AttentionMode.SHARED_PREFILL:
            if not cache.has_shared():
                out, _ = flash_attention(q, k, v, causal=True)
            else:
                out = hydragen_attention_on_caches(q, k, v, cache.get_used_shared_caches())
            cache.append_shared(k, v, ctx.shared_lens_dev, ctx.host_lens)
        elif self.mode == AttentionMode.UNIQUE_PREFILL:
            if self.disable_hydragen:
                kc, vc = cache.update_per_completion_kvs(ctx.unique_position_ids, k, v)
                n = ctx.unique_max_len
                out, _ = flash_attention(q, kc[:, :n], vc[:, :n], causal=True)
            else:
                if not cache.has_shared():
                    out, _ = flash_attention(q, k, v, causal=True)
                else:
                    out = hydragen_attention_on_caches(q, k, v, cache.get_used_shared_caches())
                cache.update_per_completion_kvs(ctx.unique_position_ids, k, v)
        elif self.mode == AttentionMode.DECODE:
            if self.fused_decode and s == 1 and self.num_key_value_groups in (1, 2, 4, 8):
                # THE HOT PATH: prefix launch(es) + one launch for append + suffix + combine
                shared = [] if (self.disable_hydragen or not cache.has_shared()) else cache.get_used_shared_caches()
                out = hydragen_decode_on_caches(q, k, v, ctx.unique_position_ids, cache, shared)
            else:  # the reference's call sequence, primitive by primitive (hydragen/llama.py:564-587)
                kc, vc = cache.update_per_completion_kvs(ctx.unique_position_ids, k, v)
                if not cache.has_shared() or self.disable_hydragen:
                    out, _ = flash_attention_seqlen(q, kc, vc, seq_len=ctx.seq_lens)
                else:
                    out = hydragen_attention_on_caches(q, kc, vc, cache.get_used_shared_caches(), seq_len=ctx.seq_lens)
        else:
            raise ValueError(f"Unknown mode {self.mode}")

        out = out.reshape(b, s, self.num_heads * self.head_dim)
        if self.all_reduce is not None:
            return _rowwise_then_all_reduce(out, self.o_proj, self.all_reduce)
        return self.o_proj(out)


class HydragenLlamaDecoderLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self_attn = HydragenLlamaAttention(config)
        self.mlp = LlamaMLP(config)
        self.input_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        self.post_attention_layernorm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)

    def forward(self, hidden_states: Tensor, ctx: StepContext) -> Tensor:
        hidden_states = hidden_states + self.self_attn(self.input_layernorm(hidden_states), ctx)
        return hidden_states + self.mlp(self.post_attention_layernorm(hidden_states))


class HydragenLlamaModel(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        pad = getattr(config, "pad_token_id", None)
        self.padding_idx = pad if pad is not None else 0
        self.vocab_size = config.vocab_size
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size, self.padding_idx)
        self.layers = nn.ModuleList([HydragenLlamaDecoderLayer(config) for _ in range(config.num_hidden_layers)])
        self.norm = LlamaRMSNorm(config.hidden_size, eps=config.rms_norm_eps)
        scaling = getattr(config, "rope_scaling", None)
        factor = 1.0
        if scaling:
            kind = scaling.get("type", scaling.get("rope_type"))
            if kind == "linear":
                factor = float(scaling["factor"])
            elif kind not in (None, "default"):
                raise ValueError(f"Unknown RoPE scaling type {kind}")
        self.rotary_emb = HydragenLlamaRotaryEmbedding(_cfg_head_dim(config), config.max_position_embeddings,
                                                       getattr(config, "rope_theta", 10000.0), factor)
        for layer in self.layers:
            # shared, and NOT registered as a submodule of every attention layer (the reference assigns a plain attribute too)
            object.__setattr__(layer.self_attn, "rotary_emb", self.rotary_emb)
        self.mode: Optional[str] = None

    # -- switches and graph-validity probes (hydragen/llama.py:666-714) ---------------------------
    def _attn(self, i: int = 0) -> HydragenLlamaAttention:
        return self.layers[i].self_attn

    def set_disable_hydragen(self, disable: bool = True):
        for layer in self.layers:
            layer.self_attn.disable_hydragen = disable

    def get_disable_hydragen(self) -> bool:
        return self._attn().disable_hydragen

    def set_disable_attention(self, disable: bool = True):
        for layer in self.layers:
            layer.self_attn.disable_attention = disable

    def get_disable_attention(self) -> bool:
        return self._attn().disable_attention

    def copy_shared_cache_to_unique(self, total_num_sequences: int):
        for layer in self.layers:
            layer.self_attn.kv_cache.copy_shared_to_unique(total_num_sequences)

    def get_shared_batch_sizes(self) -> List[int]:
        return [c.get_current_batch_size() for c in self._attn().kv_cache.get_used_shared_caches()]

    def get_shared_varlens(self) -> List[bool]:
        return [c.use_varlen for c in self._attn().kv_cache.get_used_shared_caches()]

    def get_shared_slice_seq_lens(self) -> List[Optional[int]]:
        return [c.sliced_sequence_length for c in self._attn().kv_cache.get_used_shared_caches()]

    # -- forward -----------------------------------------------------------------------------------
    def make_context(self, position_ids: Tensor, dtype: torch.dtype, valid_lens: Optional[Tensor] = None) -> StepContext:
        cache = self._attn().kv_cache
        cos, sin = self.rotary_emb.tables(dtype, position_ids.device)
        if self.get_disable_hydragen():
            upos = position_ids
        else:
            upos = position_ids - cache.get_shared_len(position_ids.shape[0]).unsqueeze(-1)
        ctx = StepContext(cos=cos, sin=sin, position_ids=position_ids, unique_position_ids=upos, seq_lens=None)
        if self.mode == AttentionMode.DECODE:
            ctx.seq_lens = upos[:, -1] + 1
        elif self.mode == AttentionMode.SHARED_PREFILL:
            # valid length of each new shared sequence: given by the caller, or (no padding) the width
            if valid_lens is None:
                ctx.host_lens = [position_ids.shape[1]] * position_ids.shape[0]
                ctx.shared_lens_dev = torch.full((position_ids.shape[0],), position_ids.shape[1], dtype=torch.long, device=position_ids.device)
            else:
                ctx.host_lens = [int(x) for x in valid_lens.tolist()]
                ctx.shared_lens_dev = valid_lens.to(position_ids.device)
        elif self.mode == AttentionMode.UNIQUE_PREFILL and self.get_disable_hydragen():
            ctx.unique_max_len = int(upos.max().item()) + 1
        return ctx

    def forward(self, input_ids: Tensor, position_ids: Tensor, valid_lens: Optional[Tensor] = None) -> Tensor:
        h = self.embed_tokens(input_ids)
        ctx = self.make_context(position_ids, h.dtype, valid_lens)
        for layer in self.layers:
            h = layer(h, ctx)
        return self.norm(h)


@dataclass
class CaptureData:
    graph: "torch.cuda.CUDAGraph"
    static_input_ids: Tensor
    static_position_ids: Tensor
    static_hidden: Tensor
    key: tuple = field(default_factory=tuple)


class GraphedHydragenLlamaModel(nn.Module):
    """CUDA-graph replay of the whole decoder stack for decode steps (hydragen/llama.py:781-866).
    The kernels behind the C ABI enqueue on the capturing stream, never allocate and never sync,
    so one decode step is a single graph launch.  Re-captured when anything baked into the graph
    changes: shapes, shared batch sizes / lengths / raggedness, the disable switches."""

    def __init__(self, model: HydragenLlamaModel):
        super().__init__()
        self.model = model
        self.capture_data: Optional[CaptureData] = None

    def invalidate(self):
        self.capture_data = None

    def _key(self, input_ids: Tensor, position_ids: Tensor) -> tuple:
        m = self.model
        return (tuple(input_ids.shape), tuple(position_ids.shape), tuple(m.get_shared_batch_sizes()), tuple(m.get_shared_varlens()),
                tuple(m.get_shared_slice_seq_lens()), m.get_disable_hydragen(), m.get_disable_attention(), m.mode)

    def forward(self, input_ids: Tensor, position_ids: Tensor) -> Tensor:
        key = self._key(input_ids, position_ids)
        if self.capture_data is not None and self.capture_data.key != key:
            self.invalidate()
        if self.capture_data is None:
            self.capture(input_ids, position_ids)
        cd = self.capture_data
        cd.static_input_ids.copy_(input_ids)
        cd.static_position_ids.copy_(position_ids)
        cd.graph.replay()
        return cd.static_hidden

    def capture(self, input_ids: Tensor, position_ids: Tensor):
        static_ids, static_pos = input_ids.clone(), position_ids.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):  # warm-up: lazy inits (cuBLAS handles, hg_init, NCCL) happen outside the capture
                self.model(input_ids=static_ids, position_ids=static_pos)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            hidden = self.model(input_ids=static_ids, position_ids=static_pos)
        self.capture_data = CaptureData(g, static_ids, static_pos, hidden, self._key(input_ids, position_ids))


class SharedCacheOp:
    WIPE = "wipe"
    EXTEND = "extend"
    PRESERVE = "preserve"


class HydragenLlamaForCausalLM(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.model = HydragenLlamaModel(config)
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.kv_cache_allocated = False
        self.graphed_model: Optional[GraphedHydragenLlamaModel] = None
        self.mode: Optional[str] = None
        self.unique_kv_storage: Optional[Tensor] = None

    # -- construction --------------------------------------------------------------------------
    @classmethod
    def from_config(cls, config, dtype: torch.dtype = torch.bfloat16, device: Union[str, torch.device] = "cuda", seed: Optional[int] = 0,
                    init_std: float = 0.02):
        """Random-init model of a given architecture, built directly on ``device`` (the container has
        no checkpoints; hydragen/llama.py:1411-1416 uses accelerate's meta init + HF weights)."""
        if isinstance(config, str):
            config = llama_config(config)
        with torch.device(device):
            model = cls(config).to(dtype)
        if seed is not None:
            g = torch.Generator(device=device).manual_seed(seed)
            with torch.no_grad():
                for name, p in model.named_parameters():
                    if p.ndim >= 2:
                        p.normal_(0.0, init_std, generator=g)
        model.device, model.dtype = torch.device(device), dtype
        return model

    @classmethod
    def from_pretrained(cls, model_name_or_path: str, **kwargs):
        """hydragen/llama.py:1398-1422: load a HuggingFace Llama checkpoint into this module tree
        (same state-dict keys).  Needs transformers and the weights on disk."""
        from transformers import LlamaForCausalLM  # imported lazily: plumbing, and slow to import

        hf_model = LlamaForCausalLM.from_pretrained(model_name_or_path, **kwargs)
        if hf_model.dtype not in (torch.float16, torch.bfloat16):
            raise ValueError(f"Model must be in float16 or bfloat16, not {hf_model.dtype}")
        with torch.device("meta"):  # parameters only: the module tree owns no buffers (see HydragenLlamaRotaryEmbedding)
            model = cls(hf_model.config)
        sd = {k: v for k, v in hf_model.state_dict().items() if "rotary_emb" not in k}
        missing, unexpected = model.load_state_dict(sd, assign=True, strict=False)
        if missing:
            raise ValueError(f"checkpoint lacks parameters of the module tree: {missing[:5]}{' ...' if len(missing) > 5 else ''}")
        model.to(hf_model.device)
        model.device, model.dtype = hf_model.device, hf_model.dtype
        return model

    # -- mode / graph ----------------------------------------------------------------------------
    def set_mode(self, mode):
        self.mode = mode
        self.model.mode = mode
        for layer in self.model.layers:
            layer.self_attn.mode = mode

    def graph(self, do_graph: bool = True):
        """Controls whether decoding replays a CUDA graph."""
        if do_graph:
            if self.graphed_model is None:
                self.graphed_model = GraphedHydragenLlamaModel(self.model)
        else:
            self.graphed_model = None

    def maybe_invalidate(self):
        if self.graphed_model is not None:
            self.graphed_model.invalidate()

    def get_num_heads(self) -> int:
        for name in ("num_heads", "num_attention_heads"):
            if hasattr(self.config, name):
                return getattr(self.config, name)
        raise ValueError("config needs to specify num heads")

    def setup_caches(self, max_unique_batch_size: int, max_unique_seq_length: int, max_shared_batch_sizes: List[int],
                     max_shared_seq_lengths: List[int]):
        """Allocates the KV caches of every layer (hydragen/llama.py:921-955).  The unique length is
        rounded up to a multiple of 16.  The unique caches of all layers live in one tensor
        [layers, 2, B, Lu, Hkv, d]."""
        self.maybe_invalidate()
        max_unique_seq_length = (max_unique_seq_length + 15) // 16 * 16
        w = self.lm_head.weight
        n_layers = len(self.model.layers)
        hkv, hd = self.config.num_key_value_heads, _cfg_head_dim(self.config)
        self.unique_kv_storage = torch.zeros((n_layers, 2, max_unique_batch_size, max_unique_seq_length, hkv, hd), dtype=w.dtype, device=w.device)
        for i, layer in enumerate(self.model.layers):
            layer.self_attn.kv_cache = PerLayerKVCache(
                max_unique_batch_size=max_unique_batch_size, max_unique_seq_length=max_unique_seq_length,
                max_shared_batch_sizes=max_shared_batch_sizes, max_shared_seq_lengths=max_shared_seq_lengths,
                n_kv_heads=hkv, head_dim=hd, device=w.device, dtype=w.dtype, storage=self.unique_kv_storage[i])
        self.kv_cache_allocated = True

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, input_ids: Tensor, position_ids: Tensor, seq_lens: Optional[Tensor] = None, use_graph: bool = False,
                full_logits: bool = False) -> Tensor:
        if use_graph:
            assert self.graphed_model is not None
            hidden = self.graphed_model(input_ids=input_ids, position_ids=position_ids)
        else:
            hidden = self.model(input_ids=input_ids, position_ids=position_ids, valid_lens=seq_lens)
        # the LM head runs on the last valid token only unless full logits are asked for
        if full_logits:
            to_head = hidden
        elif seq_lens is not None:
            idx = (seq_lens.to(hidden.device).long() - 1).view(-1, 1, 1).expand(-1, 1, hidden.shape[-1])
            to_head = hidden.gather(1, idx)
        else:
            to_head = hidden[:, -1:]
        return self.lm_head(to_head).float()

    # -- sampling (plumbing; hydragen/llama.py:999-1046) -----------------------------------------
    def apply_top_p(self, logits: Tensor, top_p: float, min_tokens_to_keep: int = 1, filter_value: float = -float("Inf")) -> Tensor:
        vals, order = torch.sort(logits, descending=False)
        drop = vals.softmax(dim=-1).cumsum(dim=-1) <= (1 - top_p)
        drop[..., -min_tokens_to_keep:] = False
        return logits.masked_fill(drop.scatter(1, order, drop), filter_value)

    def sample_from_logits(self, logits: Tensor, temperature: float, num_samples: int = 1, top_p: Optional[float] = None) -> Tensor:
        if top_p is not None:
            logits = self.apply_top_p(logits, top_p)
        if temperature == 0:
            assert logits.ndim == 2
            return logits.argmax(dim=-1, keepdim=True).repeat_interleave(num_samples, dim=-1)
        probs = nn.functional.softmax(logits / temperature, dim=-1)
        return torch.multinomial(probs, num_samples=num_samples, replacement=True)

    # -- shared-cache management -----------------------------------------------------------------
    def empty_shared_cache(self):
        for layer in self.model.layers:
            layer.self_attn.kv_cache.empty_shared_cache()

    def truncate_shared_caches(self, new_num_shared_caches: int):
        """Keeps the first ``new_num_shared_caches`` shared levels (0 removes all)."""
        for layer in self.model.layers:
            layer.self_attn.kv_cache.truncate_shared_caches(new_num_shared_caches)

    def get_shared_cache_len(self, batch_size: int) -> Tensor:
        return self.model._attn().kv_cache.get_shared_len(batch_size)

    def get_num_used_shared_caches(self) -> int:
        return self.model._attn().kv_cache.num_used_shared_caches

    def _prefill_positions(self, input_ids: Tensor, seq_lens: Optional[Tensor]) -> Tensor:
        """Absolute positions of a new block of tokens: each row continues after the shared length
        its sequence already sees; right-padding repeats the last valid position
        (hydragen/llama.py:1088-1107)."""
        bs, width = input_ids.shape
        start = self.get_shared_cache_len(bs)  # [bs] int64 (levels are repeat-interleaved to bs)
        pos = start[:, None] + torch.arange(width, device=input_ids.device, dtype=torch.long)[None, :]
        if seq_lens is not None:
            last = start + seq_lens.to(start.device).long() - 1
            pos = torch.minimum(pos, last[:, None])
        return pos

    @torch.no_grad()
    def append_shared(self, input_ids: Tensor, seq_lens: Optional[Tensor] = None, full_logits: Optional[bool] = False) -> Tensor:
        """Adds a new level of shared cache: prefill of ``input_ids`` [sb, L] (right padded,
        ``seq_lens`` = true lengths or None) attending to the existing levels."""
        self.set_mode(AttentionMode.SHARED_PREFILL)
        pos = self._prefill_positions(input_ids, seq_lens)
        return self(input_ids=input_ids, position_ids=pos, seq_lens=seq_lens, full_logits=bool(full_logits))

    @torch.no_grad()
    def process_unique(self, input_ids: Tensor, seq_lens: Optional[Tensor] = None) -> Tensor:
        """Prefill of per-sequence (non-shared) prompt tokens into the unique cache."""
        self.set_mode(AttentionMode.UNIQUE_PREFILL)
        bs, width = input_ids.shape
        cache = self.model._attn().kv_cache
        if bs > cache.per_completion_k_cache.shape[0] or width > cache.per_completion_k_cache.shape[1]:
            raise ValueError(f"[{bs}, {width}] unique tokens exceed the unique cache {tuple(cache.per_completion_k_cache.shape[:2])} (setup_caches)")
        start = self.get_shared_cache_len(bs)
        pos = start[:, None] + torch.arange(width, device=input_ids.device, dtype=torch.long)[None, :]
        return self(input_ids=input_ids, position_ids=pos, seq_lens=seq_lens)

    def repeat_per_completion_cache_for_num_samples(self, current_size: int, num_samples: int):
        for layer in self.model.layers:
            layer.self_attn.kv_cache.repeat_per_completion_cache_for_num_samples(current_size, num_samples)

    def _check_capacity(self, batch: int, shared: List[Tensor], suffix: Optional[Tensor], max_new_tokens: int, disable_hydragen: bool):
        """Fail loudly BEFORE anything is written when the request does not fit the caches or the RoPE tables (the
        kernels clamp out-of-range rows instead of faulting; the reference's scatter_ / gather would raise).  Host
        arithmetic on the padded widths -- upper bounds of the true lengths -- so there is no device read."""
        cache = self.model._attn().kv_cache
        cap_b, cap_len = cache.per_completion_k_cache.shape[0], cache.per_completion_k_cache.shape[1]
        if batch > cap_b:
            raise ValueError(f"{batch} sequences exceed max_unique_batch_size = {cap_b} (setup_caches)")
        shared_len = sum(c.max_used_length for c in cache.get_used_shared_caches()) + sum(int(t.shape[1]) for t in shared)
        suffix_w = int(suffix.shape[1]) if suffix is not None else 0
        unique_rows = (shared_len if disable_hydragen else 0) + suffix_w + max(0, max_new_tokens - 1)
        if unique_rows > cap_len:
            raise ValueError(f"{unique_rows} tokens per sequence ({suffix_w} prompt + {max_new_tokens} new"
                             f"{' + ' + str(shared_len) + ' copied shared' if disable_hydragen else ''}) exceed max_unique_seq_length = {cap_len} (setup_caches)")
        max_pos = int(self.config.max_position_embeddings)
        if shared_len + suffix_w + max(0, max_new_tokens - 1) > max_pos:
            raise ValueError(f"positions up to {shared_len + suffix_w + max_new_tokens - 2} exceed max_position_embeddings = {max_pos} (RoPE table rows)")

    # -- generation ------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(
        self,
        input_ids: Optional[Union[Tensor, List[Tensor]]] = None,
        seq_lens: Optional[Union[Tensor, List[Tensor]]] = None,
        starting_logits: Optional[Tensor] = None,
        num_return_sequences: int = 1,
        max_new_tokens: int = 5,
        temperature: float = 1.0,
        top_p: Optional[float] = None,
        eos_token_id: Optional[int] = None,
        return_logits: bool = False,
        shared_cache_op: str = SharedCacheOp.PRESERVE,
        disable_hydragen: bool = False,
        disable_attention: bool = False,
        disable_hierarchy: bool = False,
        token_overrides: Optional[Tensor] = None,
    ):
        """Same contract as hydragen/llama.py:1157-1396.

        ``input_ids``: one id tensor [batch, len] or a list of them forming a prompt hierarchy (every
        batch size divides the last one; right padded, with ``seq_lens`` giving true lengths).  With
        ``num_return_sequences > 1`` every given tensor is a SHARED level and the completions form
        the last level; otherwise the last tensor is processed into the per-sequence cache.
        ``starting_logits`` [batch, vocab] replaces ``input_ids`` to continue from cached prefixes.
        ``shared_cache_op``: "wipe" clears shared levels first, "preserve" (default) drops the
        levels this call added afterwards, "extend" keeps them.  ``token_overrides``
        [batch, max_new_tokens] teacher-forces the fed-back tokens (parity tests).
        ``disable_hydragen`` / ``disable_attention`` / ``disable_hierarchy`` are the benchmarking
        baselines of the reference.  Returns ids [batch * num_return_sequences, <= max_new_tokens]
        (and the list of per-step logits when ``return_logits``).
        """
        assert self.kv_cache_allocated
        assert (input_ids is None) != (starting_logits is None), "give exactly one of input_ids / starting_logits"
        if temperature < 0:
            raise ValueError(f"temperature must be non-negative, {temperature} is invalid")
        levels: List[Tensor] = [] if input_ids is None else ([input_ids] if isinstance(input_ids, Tensor) else list(input_ids))
        if isinstance(seq_lens, Tensor):
            level_lens: List[Optional[Tensor]] = [seq_lens]
        elif seq_lens is None:
            level_lens = [None] * len(levels)
        else:
            level_lens = list(seq_lens)
        assert len(level_lens) == len(levels)

        if disable_attention:
            self.model.set_disable_attention(True)
        if shared_cache_op == SharedCacheOp.WIPE:
            self.empty_shared_cache()
        og_levels = self.get_num_used_shared_caches()
        nrs = num_return_sequences
        new_levels = len(levels) + (1 if nrs > 1 else 0)
        total_levels = og_levels + new_levels
        if disable_hydragen:  # FlashAttention baseline: prefix + completions, or prefix + suffix + 1 completion
            assert total_levels == 2
            if new_levels == 2:
                assert levels[0].shape[0] == 1
        if disable_hierarchy:  # single-level Hydragen baseline: prefix + suffix + many completions
            assert total_levels == 3 and nrs > 1

        batch = (levels[-1].shape[0] if levels else starting_logits.shape[0]) * nrs
        all_shared = nrs > 1 and not (disable_hierarchy or disable_hydragen)
        if all_shared or not levels:
            shared, shared_lens, suffix, suffix_lens = levels, level_lens, None, None
        else:
            shared, shared_lens, suffix, suffix_lens = levels[:-1], level_lens[:-1], levels[-1], level_lens[-1]
        self._check_capacity(batch, shared, suffix, max_new_tokens, disable_hydragen)

        logits = None if starting_logits is None else starting_logits.unsqueeze(1)
        for ids, lens in zip(shared, shared_lens):
            logits = self.append_shared(ids, lens)
        if disable_hydragen:
            self.model.set_disable_hydragen(True)
            if self.get_num_used_shared_caches() > 0:
                self.model.copy_shared_cache_to_unique(batch)
        if suffix is not None:
            logits = self.process_unique(suffix, suffix_lens)
            self.repeat_per_completion_cache_for_num_samples(suffix.shape[0], nrs)

        prefill_logits = logits[:, -1]
        first = self.sample_from_logits(prefill_logits, temperature=temperature, num_samples=nrs, top_p=top_p).reshape(-1, 1)
        step_logits = [prefill_logits.repeat_interleave(nrs, 0)] if return_logits else None

        start_pos = self.get_shared_cache_len(first.shape[0])[:, None]
        if suffix is not None:
            sl = suffix_lens if suffix_lens is not None else torch.full((suffix.shape[0],), suffix.shape[1], dtype=torch.long, device=suffix.device)
            start_pos = start_pos + sl.to(start_pos.device).long().repeat_interleave(nrs, 0)[:, None]

        finished = (first == eos_token_id) if eos_token_id is not None else None
        decoded = [first]
        current = first if token_overrides is None else token_overrides[:, 0:1]

        self.set_mode(AttentionMode.DECODE)
        use_graph = self.graphed_model is not None
        for t in range(max_new_tokens - 1):
            out = self(input_ids=current, position_ids=start_pos + t, use_graph=use_graph)
            if return_logits:
                step_logits.append(out[:, -1])
            current = self.sample_from_logits(out[:, -1], temperature=temperature, top_p=top_p)
            if finished is not None:
                finished = torch.logical_or(finished, current == eos_token_id)
                if torch.all(finished):
                    break
            decoded.append(current)
            if token_overrides is not None:
                current = token_overrides[:, t + 1 : t + 2]

        ids_out = torch.cat(decoded, dim=-1)
        if shared_cache_op == SharedCacheOp.PRESERVE:
            self.truncate_shared_caches(og_levels)
        if disable_hydragen:
            self.model.set_disable_hydragen(False)
        if disable_attention:
            self.model.set_disable_attention(False)
        return (ids_out, step_logits) if return_logits else ids_out
