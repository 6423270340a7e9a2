"""The collective that follows the attention hot path under head-axis tensor parallelism: all-reduce(sum) of the
row-parallel o_proj output, one per attention layer (hydragen/tp.py:108-112).

``MultimemAllReduce`` runs it as the library's own NVLS kernel (csrc/allreduce.cu: multimem.ld_reduce / multimem.st
through the NVSwitch) on buffers that live in symmetric memory; torch supplies the plumbing (symmetric allocation,
rendezvous, multicast mapping: ``torch.distributed._symmetric_memory``).  Where the platform has no multicast
support the caller keeps using ``torch.distributed.all_reduce`` (NCCL) -- ``MultimemAllReduce.available`` says which.
"""

from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _lib


GEMM_FLAG_WORDS = 16384  # enough for 2032 output tiles of 128 x 256 at 8 ranks


class MultimemAllReduce:
    def __init__(self, nbytes: int, device: torch.device, group=None, n_blocks: int = 16):
        """Collective over ``group`` (default: world) for messages carved out of one symmetric arena of ``nbytes``."""
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = device
        self.n_blocks = n_blocks
        name = self.group.group_name
        self.arena = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self._h = symm_mem.rendezvous(self.arena, name)
        # flag words of csrc/allreduce.cu: [0] epoch, [1] CTA count, [32 + p] / [64 + p] per-peer arrival epochs
        self.flags = symm_mem.empty(128, dtype=torch.int32, device=device)
        self.flags.zero_()
        self._hf = symm_mem.rendezvous(self.flags, name)
        # flag words of csrc/oproj_allreduce.cu (the GEMM fused with this collective): its own array -- per-tile arrival
        # epochs, 128 + tiles * world words
        self.gemm_flags = symm_mem.empty(GEMM_FLAG_WORDS, dtype=torch.int32, device=device)
        self.gemm_flags.zero_()
        self._hg = symm_mem.rendezvous(self.gemm_flags, name)
        self.mc_base = int(self._h.multicast_ptr or 0)  # 0: no multicast mapping on this platform
        self.flags_dev = int(self._hf.buffer_ptrs_dev)
        self.gemm_flags_dev = int(self._hg.buffer_ptrs_dev)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)
        self._used = 0

    @property
    def available(self) -> bool:
        return self.mc_base != 0

    def buffer(self, shape, dtype: torch.dtype) -> torch.Tensor:
        """A tensor inside the symmetric arena (same offset on every rank, 256-byte aligned)."""
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        off = (self._used + 255) // 256 * 256
        if off + nbytes > self.arena.numel():
            raise ValueError("symmetric arena exhausted")
        self._used = off + nbytes
        return self.arena[off : off + nbytes].view(dtype).view(*shape)

    def all_reduce_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over the group of a tensor obtained from ``buffer()`` (contiguous)."""
        off = t.data_ptr() - self.arena.data_ptr()
        nbytes = t.numel() * t.element_size()
        assert 0 <= off and off + nbytes <= self.arena.numel() and t.is_contiguous(), "tensor must come from buffer()"
        assert self.available, "no NVLink multicast mapping on this platform: use torch.distributed.all_reduce"
        _lib.allreduce_multimem(self.mc_base + off, 0, self.flags_dev, self.rank, self.world, nbytes, t.dtype, self.n_blocks, self.device)
        return t

    def linear_all_reduce_(self, x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, n_ctas: int = 0) -> torch.Tensor:
        """out = sum over the group of x @ w.T in ONE launch (csrc/oproj_allreduce.cu): x [m, k] this rank's attention
        output, w [n, k] its slice of o_proj.weight, out [m, n] from ``buffer()`` -- hydragen/llama.py:592-594 followed by
        hydragen/tp.py:108-112."""
        off = out.data_ptr() - self.arena.data_ptr()
        nbytes = out.numel() * out.element_size()
        assert 0 <= off and off + nbytes <= self.arena.numel() and out.is_contiguous(), "out must come from buffer()"
        assert self.available, "no NVLink multicast mapping on this platform"
        if _lib.oproj_allreduce_flag_words(x.shape[0], w.shape[0], self.world) > GEMM_FLAG_WORDS:
            raise ValueError(f"linear_all_reduce_: [{x.shape[0]}, {w.shape[0]}] has too many tiles for {GEMM_FLAG_WORDS} flag words")
        _lib.oproj_allreduce_fwd(x, w, out, self.mc_base + off, self.gemm_flags_dev, GEMM_FLAG_WORDS, self.rank, self.world, n_ctas)
        return out

    def all_reduce(self, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Out-of-place sum (one-shot form: one cross-rank barrier): ``t`` from ``buffer()``, result in a private
        tensor.  ``t`` may be overwritten once a later collective of this object has completed."""
        off = t.data_ptr() - self.arena.data_ptr()
        nbytes = t.numel() * t.element_size()
        assert 0 <= off and off + nbytes <= self.arena.numel() and t.is_contiguous(), "tensor must come from buffer()"
        assert self.available, "no NVLink multicast mapping on this platform: use torch.distributed.all_reduce"
        if out is None:
            out = torch.empty_like(t)
        _lib.allreduce_multimem(self.mc_base + off, out.data_ptr(), self.flags_dev, self.rank, self.world, nbytes, t.dtype, self.n_blocks, self.device)
        return out


def make_all_reduce(nbytes: int, device: torch.device, group=None) -> Optional[MultimemAllReduce]:
    """The NVLS collective if EVERY rank of the group could set it up, else None on every rank (the caller then uses
    NCCL).  Collective: all ranks must call it together.  The ranks agree on the outcome -- a rank that fell back on
    its own would leave the others spinning in the kernel's cross-rank flags."""
    ar, ok = None, 0
    try:
        ar = MultimemAllReduce(nbytes, device, group)
        ok = 1 if ar.available else 0
    except Exception as ex:
        import warnings

        warnings.warn(f"NVLS all-reduce unavailable on this rank ({ex!r}); the group will agree on NCCL")
    flag = torch.tensor([ok], device=device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return ar if int(flag.item()) == 1 else None
