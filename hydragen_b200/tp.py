"""Head-axis tensor parallelism for the Hydragen Llama path: one process per GPU, NCCL over
NVLink / NVSwitch.  Same split as the reference's ``hydragen/tp.py`` (:30-124):

* ``q_proj / k_proj / v_proj`` column-parallel -> rank r owns query heads
  ``[r*Hq/ws, (r+1)*Hq/ws)`` and kv heads ``[r*Hkv/ws, (r+1)*Hkv/ws)``; the attention operator
  (prefix, suffix, combine) then runs on local heads with NO communication -- every
  (q head, kv head) pair is independent;
* ``o_proj`` row-parallel, followed by ONE all-reduce(sum) of ``[B, nq, hidden]`` per attention
  layer (tp.py:108-112);
* MLP ``gate/up`` column-parallel, ``down`` row-parallel + all-reduce (tp.py:73-87; not on the
  attention path but required for exact results);
* embeddings, norms, ``lm_head`` and sampling are replicated with the same RNG seed on every rank
  (tp.py:178).

The collective is ``torch.distributed.all_reduce`` issued in ``forward`` on the compute stream: it
is captured into the decode CUDA graph together with the kernels, so a decode step stays one graph
launch per rank.  ``backend="gloo"`` (CPU) runs the same code in the unit tests.
"""

from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.distributed as dist
from torch import Tensor, nn

from .llama import HydragenLlamaAttention, HydragenLlamaForCausalLM, HydragenLlamaModel, LlamaMLP
from .utils import get_rank, get_world_size

_COLWISE = ("q_proj", "k_proj", "v_proj", "gate_proj", "up_proj")
_ROWWISE = ("o_proj", "down_proj")


def _shard(x: Tensor, dim: int, rank: int, world_size: int) -> Tensor:
    assert x.size(dim) % world_size == 0, f"cannot split {tuple(x.shape)} dim {dim} over {world_size} ranks"
    n = x.size(dim) // world_size
    return x.narrow(dim, rank * n, n).clone()


def shard_state_dict(sd: Dict[str, Tensor], rank: int, world_size: int) -> Dict[str, Tensor]:
    """Rank ``rank``'s slice of a full (HuggingFace-keyed) Llama state dict -- what the reference's
    ``make_tp_files.py`` writes to ``{rank}.pt``."""
    out = {}
    for name, w in sd.items():
        leaf = name.split(".")[-2] if "." in name else name
        if leaf in _COLWISE:
            out[name] = _shard(w, 0, rank, world_size)
        elif leaf in _ROWWISE and name.endswith("weight"):
            out[name] = _shard(w, 1, rank, world_size)
        else:
            out[name] = w
    return out


_FUSED_LINEAR = os.environ.get("HYDRAGEN_B200_FUSED_LINEAR", "1") != "0"  # 0: library GEMM, then the stand-alone collective


class _AllReduce:
    """Sum over the tensor-parallel group on the compute stream.  Messages up to ``NVLS_MAX_BYTES`` (the decode-step
    ``[B, 1, hidden]`` sizes) go through the library's NVLS kernel (csrc/allreduce.cu) where the platform has NVLink
    multicast; everything else -- prefill activations, whose shape changes with every prompt -- through
    ``torch.distributed.all_reduce`` (NCCL; gloo in the CPU tests).

    ONE symmetric arena per process group, created on first use (every rank issues the same collectives in the same
    order, so they all get here together) and never grown: two halves used alternately, so a result stays valid until
    the second-next call.  Whether the NVLS path is used is AGREED across the ranks (a MIN all-reduce of "my set-up
    worked"): a rank that fell back on its own would leave the others spinning in the kernel's flag barrier."""

    NVLS_MAX_BYTES = int(os.environ.get("HYDRAGEN_B200_NVLS_MAX_MIB", "32")) << 20
    _arenas: Dict[int, object] = {}  # id(group) -> [MultimemAllReduce, [half0, half1] uint8 views, turn] or None (agreed: not available)

    def __init__(self, group=None, use_nvls: bool = True):
        self.group = group
        self.use_nvls = use_nvls

    def _arena(self, device: torch.device):
        key = id(self.group)
        if key not in _AllReduce._arenas:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the NVLS all-reduce must be set up before CUDA-graph capture (run one eager step first)")
            entry, ok = None, 0
            try:
                from .collectives import MultimemAllReduce

                ar = MultimemAllReduce(2 * self.NVLS_MAX_BYTES + 512, device, self.group)
                if ar.available:
                    entry = [ar, [ar.buffer((self.NVLS_MAX_BYTES,), torch.uint8) for _ in range(2)], 0]  # collective, halves, whose turn
                    ok = 1
            except Exception as ex:  # no symmetric memory / no multicast on this platform
                import warnings

                warnings.warn(f"NVLS all-reduce unavailable on this rank ({ex!r}); the group will agree on NCCL")
            flag = torch.tensor([ok], device=device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            _AllReduce._arenas[key] = entry if int(flag.item()) == 1 else None
        return _AllReduce._arenas[key]

    def linear(self, x: Tensor, lin: nn.Linear) -> Tensor:
        """all_reduce(lin(x)) for a row-parallel ``lin`` (o_proj: hydragen/llama.py:592-594 + hydragen/tp.py:108-112; the
        MLP's down_proj alike).  Decode-sized 16-bit products run as ONE launch -- the tcgen05 GEMM whose tiles the NVSwitch
        reduces as they are produced (csrc/oproj_allreduce.cu); everything else as the GEMM followed by ``__call__``."""
        n, k = lin.weight.shape
        m = x.numel() // max(k, 1)
        nbytes = m * n * x.element_size()
        if (self.use_nvls and _FUSED_LINEAR and x.is_cuda and lin.bias is None and x.dtype in (torch.bfloat16, torch.float16)
                and lin.weight.dtype == x.dtype and m > 0 and 0 < nbytes <= self.NVLS_MAX_BYTES and n % 8 == 0 and k % 8 == 0
                and dist.get_backend(self.group) == "nccl"):
            entry = self._arena(x.device)
            if entry is not None:
                from . import _lib
                from .collectives import GEMM_FLAG_WORDS

                ar, halves, turn = entry
                x2 = x.reshape(m, k)
                if x2.stride(1) == 1 and x2.stride(0) % 8 == 0 and x2.data_ptr() % 16 == 0 and lin.weight.is_contiguous() \
                        and _lib.oproj_allreduce_flag_words(m, n, ar.world) <= GEMM_FLAG_WORDS:
                    buf = halves[turn][:nbytes].view(x.dtype).view(m, n)
                    entry[2] = turn ^ 1
                    ar.linear_all_reduce_(x2, lin.weight, buf)
                    return buf.view(*x.shape[:-1], n)
        return self(lin(x))

    def __call__(self, x: Tensor) -> Tensor:
        nbytes = x.numel() * x.element_size()
        if (self.use_nvls and x.is_cuda and 0 < nbytes <= self.NVLS_MAX_BYTES and nbytes % 16 == 0 and x.dtype in (torch.bfloat16, torch.float16, torch.float32)
                and dist.get_backend(self.group) == "nccl"):
            entry = self._arena(x.device)
            if entry is not None:
                ar, halves, turn = entry
                buf = halves[turn][:nbytes].view(x.dtype).view(x.shape)
                entry[2] = turn ^ 1  # shared by every layer of the model: the arena, not the layer, alternates
                buf.copy_(x)
                ar.all_reduce_(buf)
                return buf
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)
        return x


def _apply_tp_linear(linear: nn.Linear, style: str, rank: int, world_size: int) -> None:
    dim, attr = {"colwise": (0, "out_features"), "rowwise": (1, "in_features")}[style]
    assert getattr(linear, attr) % world_size == 0
    linear.weight = nn.Parameter(_shard(linear.weight.data, dim, rank, world_size), requires_grad=False)
    if linear.bias is not None and style == "colwise":
        linear.bias = nn.Parameter(_shard(linear.bias.data, 0, rank, world_size), requires_grad=False)
    setattr(linear, attr, getattr(linear, attr) // world_size)


def _apply_tp_ffn(mlp: LlamaMLP, rank: int, world_size: int, group) -> None:
    _apply_tp_linear(mlp.gate_proj, "colwise", rank, world_size)
    _apply_tp_linear(mlp.up_proj, "colwise", rank, world_size)
    _apply_tp_linear(mlp.down_proj, "rowwise", rank, world_size)
    mlp.all_reduce = _AllReduce(group)


def _apply_tp_attn(attn: HydragenLlamaAttention, rank: int, world_size: int, group) -> None:
    assert attn.num_heads % world_size == 0 and attn.num_key_value_heads % world_size == 0, (
        f"heads ({attn.num_heads} q / {attn.num_key_value_heads} kv) must divide over {world_size} ranks")
    for name in ("q_proj", "k_proj", "v_proj"):
        _apply_tp_linear(getattr(attn, name), "colwise", rank, world_size)
    _apply_tp_linear(attn.o_proj, "rowwise", rank, world_size)
    attn.num_heads //= world_size
    attn.num_key_value_heads //= world_size
    attn.all_reduce = _AllReduce(group)  # the single collective of the attention hot path


def apply_tp(model: HydragenLlamaModel, rank: Optional[int] = None, world_size: Optional[int] = None, group=None) -> None:
    """Shard ``model`` (a full, replicated HydragenLlamaModel) in place for this rank.  KV caches must
    be (re)allocated afterwards: ``local_kv_heads`` on the config tells ``setup_caches`` the local
    head count.  Unlike the reference (tp.py:115-124) ``hidden_size`` is left untouched: norms and
    embeddings are replicated at full width; only head counts change."""
    rank = get_rank() if rank is None else rank
    world_size = get_world_size() if world_size is None else world_size
    if world_size == 1:
        return
    for block in model.layers:
        _apply_tp_ffn(block.mlp, rank, world_size, group)
        _apply_tp_attn(block.self_attn, rank, world_size, group)
    cfg = model.config
    if getattr(cfg, "head_dim", None) is None:
        cfg.head_dim = cfg.hidden_size // cfg.num_attention_heads
    cfg.num_attention_heads //= world_size
    cfg.num_key_value_heads //= world_size
    cfg.tp_world_size = world_size


def from_config_tp(config, dtype: torch.dtype = torch.bfloat16, device=None, seed: int = 0, group=None) -> HydragenLlamaForCausalLM:
    """Random-init tensor-parallel model: every rank builds the SAME full model from ``seed`` (layer
    by layer, so the peak is one full layer, not one full model) and keeps its shard.  Stands in
    for ``from_pretrained_tp`` (tp.py:135-180), which needs checkpoints that do not exist offline."""
    rank, world_size = get_rank(), get_world_size()
    if device is None:
        device = f"cuda:{rank}" if torch.cuda.is_available() else "cpu"
    model = HydragenLlamaForCausalLM.from_config(config, dtype=dtype, device=device, seed=seed)
    apply_tp(model.model, rank, world_size, group)
    torch.manual_seed(1234)  # identical sampling on every rank (tp.py:178)
    return model


def from_pretrained_tp(model_name: str, load_dir, dtype: Optional[torch.dtype] = None, device=None):
    """hydragen/tp.py:135-180: build the sharded module tree and load ``{load_dir}/{rank}.pt`` (the files
    ``shard_state_dict`` / the reference's ``make_tp_files.py`` produce).  ``device`` defaults to this rank's GPU."""
    from pathlib import Path

    from transformers import LlamaConfig as HFLlamaConfig

    config = HFLlamaConfig.from_pretrained(model_name)
    rank, world_size = get_rank(), get_world_size()
    if device is None:
        device = f"cuda:{rank}" if torch.cuda.is_available() else "cpu"
    parts = sorted(Path(load_dir).glob("*.pt"), key=lambda p: int(p.stem))
    assert len(parts) == world_size, f"{len(parts)} != {world_size}"
    with torch.device("meta"):  # parameters only: the module tree owns no buffers (HydragenLlamaRotaryEmbedding)
        model = HydragenLlamaForCausalLM(config)
        apply_tp(model.model, rank, world_size)
    sd = torch.load(parts[rank], map_location=device)
    missing, _ = model.load_state_dict(sd, assign=True, strict=False)
    if missing:
        raise ValueError(f"{parts[rank]} lacks parameters of the sharded module tree: {missing[:5]}{' ...' if len(missing) > 5 else ''}")
    model.to(device)
    model.device = device
    if dtype is None or dtype == "auto":
        model.dtype = next(model.parameters()).dtype
    else:
        model.dtype = dtype
        model.to(dtype=dtype)
    torch.manual_seed(1234)
    return model
