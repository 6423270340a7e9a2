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


class _AllReduce:
    """Sum over the tensor-parallel group on the compute stream.  On CUDA with NVLink multicast the library's
    NVLS kernel (csrc/allreduce.cu) on a symmetric staging buffer; otherwise ``torch.distributed.all_reduce``
    (NCCL on GPUs without multicast, gloo in the CPU tests).  Shared by every layer of a model: one staging
    buffer per message size, set up on first use (the warm-up runs before a CUDA-graph capture)."""

    _nvls: Dict[tuple, object] = {}

    def __init__(self, group=None, use_nvls: bool = True):
        self.group = group
        self.use_nvls = use_nvls

    def _staging(self, x: Tensor):
        key = (x.device, x.dtype, tuple(x.shape), id(self.group))
        if key not in _AllReduce._nvls:
            entry = None
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the NVLS all-reduce must be set up before CUDA-graph capture (run one eager step first)")
            try:
                from .collectives import make_all_reduce

                ar = make_all_reduce(x.numel() * x.element_size() + 256, x.device, self.group)
                if ar is not None:
                    entry = (ar, ar.buffer(tuple(x.shape), x.dtype))
            except Exception:
                entry = None
            _AllReduce._nvls[key] = entry
        return _AllReduce._nvls[key]

    def __call__(self, x: Tensor) -> Tensor:
        if self.use_nvls and x.is_cuda and dist.get_backend(self.group) == "nccl":
            entry = self._staging(x)
            if entry is not None:
                ar, buf = entry
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


def from_pretrained_tp(model_name: str, load_dir, dtype: Optional[torch.dtype] = None):
    """hydragen/tp.py:135-180: build the sharded module tree and load ``{load_dir}/{rank}.pt``."""
    from pathlib import Path

    from transformers import LlamaConfig as HFLlamaConfig

    config = HFLlamaConfig.from_pretrained(model_name)
    rank, world_size = get_rank(), get_world_size()
    device = f"cuda:{rank}"
    parts = sorted(Path(load_dir).glob("*.pt"))
    assert len(parts) == world_size, f"{len(parts)} != {world_size}"
    with torch.device("meta"):
        model = HydragenLlamaForCausalLM(config)
        apply_tp(model.model, rank, world_size)
    sd = torch.load(parts[rank], map_location=device)
    model.load_state_dict(sd, assign=True, strict=False)
    model.to(device)
    model.device = device
    if dtype is None or dtype == "auto":
        model.dtype = next(model.parameters()).dtype
    else:
        model.dtype = dtype
        model.to(dtype=dtype)
    torch.manual_seed(1234)
    return model
