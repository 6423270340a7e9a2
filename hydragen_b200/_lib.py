"""ctypes binding of libhydragen_b200.so (the C ABI declared in include/hydragen_b200.h).

There is NO fallback: if the shared library is missing, or a call fails, an exception is raised.
torch is used only to hold device memory and to name the current stream.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p
from typing import Optional, Sequence

import torch

from . import build as _build

HG_F16, HG_BF16, HG_F32 = 0, 1, 2
HG_MAX_COMBINE = 8
ABI_VERSION = 2
MAX_PREFIX_LEVELS = 4

_DTYPES = {torch.float16: HG_F16, torch.bfloat16: HG_BF16, torch.float32: HG_F32}

class PrefixLevel(Structure):
    """hg_prefix_level of include/hydragen_b200.h."""

    _fields_ = [("k", c_void_p), ("v", c_void_p), ("out", c_void_p), ("lse", c_void_p), ("cu_seqlens_k", c_void_p),
                ("n_k_rows", c_int64), ("kv_stride_row", c_int64), ("n_groups", c_int32), ("k_len", c_int32),
                ("max_k_len", c_int32), ("reserved", c_int32)]


# every symbol include/hydragen_b200.h declares: name -> (restype, argtypes)
_c_void_pp = POINTER(c_void_p)
SYMBOLS = {
    "hg_abi_version": (c_int, []),
    "hg_last_error": (c_char_p, []),
    "hg_init": (c_int, [c_int]),
    "hg_sm_count": (c_int, []),
    "hg_combine_lse": (c_int, [_c_void_pp, _c_void_pp, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "hg_rowwise_attn_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
         c_int, c_int, c_int, c_int, c_int, c_int,
         c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
         _c_void_pp, _c_void_pp, c_int, c_float, c_int, c_void_p],
    ),
    "hg_prefix_attn_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_int,
         c_int, c_int, c_int, c_int64, c_int64, c_float, c_int, c_void_p],
    ),
    "hg_prefix_attn_split_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_void_p, c_int,
         c_int, c_int, c_int, c_int64, c_int64, c_float, c_int, c_int, c_void_p],
    ),
    "hg_prefix_suggest_splits": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "hg_prefix_attn_grouped_fwd": (
        c_int,
        [c_void_p, c_int64, c_int64, POINTER(PrefixLevel), c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_int64, c_void_p],
    ),
    "hg_prefix_workspace_bytes": (c_int64, []),
    "hg_prefix_schedule": (c_int, [POINTER(PrefixLevel), c_int, c_int64, c_int, c_int, c_int, POINTER(c_int32), c_int, POINTER(c_int32)]),
    "hg_causal_attn_fwd": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_int64,
         c_float, c_int, c_void_p],
    ),
    "hg_decode_attn_fused": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
         c_int, c_int, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_int64,
         _c_void_pp, _c_void_pp, c_int, c_float, c_int, c_void_p],
    ),
    "hg_allreduce_multimem": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int, c_int, c_void_p]),
    "hg_oproj_allreduce_fwd": (
        c_int,
        [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, c_int64, c_int64, c_int,
         c_int, c_void_p],
    ),
    "hg_oproj_allreduce_flag_words": (c_int, [c_int64, c_int64, c_int]),
    "hg_oproj_allreduce_plan": (c_int, [c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hg_kv_append": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    ),
    "hg_rope_qk": (
        c_int,
        [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int,
         c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p],
    ),
}

_lib: Optional[ctypes.CDLL] = None
_inited_devices: set = set()
_prefix_ws: dict = {}  # device index -> zero-initialised workspace of the persistent prefix kernel


class HydragenB200Error(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    """dlopen the in-tree library and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise HydragenB200Error(
            f"{path} not found: the CUDA extension has not been built. Run `python -m hydragen_b200.build` "
            "(or __graft_entry__.build()). There is no CPU or PyTorch fallback for this path."
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.hg_abi_version() != ABI_VERSION:
        raise HydragenB200Error(f"ABI mismatch: library {lib.hg_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().hg_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise HydragenB200Error(f"{what} failed ({rc}): {msg}")


def ensure_init(device: torch.device):
    if device.type != "cuda":
        raise HydragenB200Error(f"hydragen_b200 kernels need CUDA tensors (got device {device}); there is no CPU path")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _inited_devices:
        _check(load().hg_init(idx), "hg_init")
        _inited_devices.add(idx)
    return idx


def prefix_workspace(device: torch.device) -> torch.Tensor:
    """The per-device workspace of the persistent prefix kernel (stream-K partials + flags): allocated and zeroed
    once, then owned by the library's launches.  All prefix launches of a process on one device share it, so they must
    be stream-ordered (they are: every caller in this package launches on the current stream).  Allocate your own
    (``torch.zeros(hg_prefix_workspace_bytes)``) for launches that may overlap on different streams."""
    idx = ensure_init(device)
    ws = _prefix_ws.get(idx)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            raise HydragenB200Error("the prefix workspace must exist before CUDA-graph capture: run one eager step first")
        ws = torch.zeros(int(load().hg_prefix_workspace_bytes()), dtype=torch.uint8, device=torch.device("cuda", idx))
        _prefix_ws[idx] = ws
    return ws


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise ValueError(f"unsupported dtype {dt}") from None


def _stream(t: torch.Tensor) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def _ptr_table(ts: Sequence[torch.Tensor]):
    arr = (c_void_p * max(1, len(ts)))()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return ctypes.cast(arr, _c_void_pp)


# ------------------------------------------------------------------------------------------
# thin typed wrappers (one per C entry point)
# ------------------------------------------------------------------------------------------


def combine_lse(outs: Sequence[torch.Tensor], lses: Sequence[torch.Tensor], out: torch.Tensor,
                lse_out: Optional[torch.Tensor]) -> None:
    ensure_init(out.device)
    d = out.shape[-1]
    rows = out.numel() // d if d > 0 else 0
    with torch.cuda.device(out.device):
        rc = load().hg_combine_lse(_ptr_table(outs), _ptr_table(lses), len(outs), _ptr(out), _ptr(lse_out), rows, d,
                                   dtype_code(out.dtype), _stream(out))
    _check(rc, "hg_combine_lse")


def rowwise_attn_fwd(q, k, v, seq_lens, cu_seqlens_k, kv_group_size, causal, out, lse, lk,
                     kv_strides, partial_outs, partial_lses, sm_scale) -> None:
    ensure_init(q.device)
    b, nq, hq, d = q.shape
    hkv = k.shape[-2]
    sl_i64 = 0
    if seq_lens is not None:
        if seq_lens.dtype == torch.int64:
            sl_i64 = 1
        elif seq_lens.dtype != torch.int32:
            raise ValueError(f"seq_lens must be int32 or int64, got {seq_lens.dtype}")
    with torch.cuda.device(q.device):
        rc = load().hg_rowwise_attn_fwd(
            _ptr(q), _ptr(k), _ptr(v), _ptr(seq_lens), sl_i64, _ptr(cu_seqlens_k), kv_group_size, int(bool(causal)),
            _ptr(out), _ptr(lse), b, nq, lk, hq, hkv, d,
            q.stride(0), q.stride(1), q.stride(2), kv_strides[0], kv_strides[1], kv_strides[2],
            _ptr_table(partial_outs), _ptr_table(partial_lses), len(partial_outs), float(sm_scale),
            dtype_code(q.dtype), _stream(q))
    _check(rc, "hg_rowwise_attn_fwd")


def prefix_attn_fwd(q, k, v, out, lse, n_groups, q_per_group, n_k_rows, k_len, cu_seqlens_k, max_k_len,
                    hq, hkv, d, q_stride_row, kv_stride_row, sm_scale, kv_splits: int = 1) -> None:
    """One shared level on the one-CTA-per-unit kernel.  out / lse hold kv_splits partial results back to back
    ([kv_splits, rows, hq, d] / [kv_splits, rows, hq])."""
    ensure_init(q.device)
    with torch.cuda.device(q.device):
        rc = load().hg_prefix_attn_split_fwd(
            _ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(lse), n_groups, q_per_group, n_k_rows, k_len,
            _ptr(cu_seqlens_k), max_k_len, hq, hkv, d, q_stride_row, kv_stride_row, float(sm_scale),
            dtype_code(q.dtype), kv_splits, _stream(q))
    _check(rc, "hg_prefix_attn_fwd")


def prefix_suggest_splits(device, n_groups: int, q_per_group: int, hq: int, max_k_len: int, max_splits: int) -> int:
    ensure_init(device)
    return int(load().hg_prefix_suggest_splits(n_groups, q_per_group, hq, max_k_len, max_splits))


def make_prefix_level(k, v, out, lse, cu_seqlens_k, n_k_rows, kv_stride_row, n_groups, k_len, max_k_len) -> PrefixLevel:
    return PrefixLevel(k.data_ptr(), v.data_ptr(), out.data_ptr(), 0 if lse is None else lse.data_ptr(),
                       0 if cu_seqlens_k is None else cu_seqlens_k.data_ptr(), n_k_rows, kv_stride_row, n_groups, k_len, max_k_len, 0)


def prefix_attn_grouped_fwd(q, n_q_rows, q_stride_row, levels: Sequence[PrefixLevel], hq, hkv, d, sm_scale,
                            workspace: Optional[torch.Tensor] = None, split: bool = True) -> None:
    """Every shared level of a hierarchy in one persistent launch (hg_prefix_attn_grouped_fwd)."""
    ensure_init(q.device)
    ws = (workspace if workspace is not None else prefix_workspace(q.device)) if split else None
    arr = (PrefixLevel * len(levels))(*levels)
    with torch.cuda.device(q.device):
        rc = load().hg_prefix_attn_grouped_fwd(_ptr(q), n_q_rows, q_stride_row, arr, len(levels), hq, hkv, d, float(sm_scale),
                                               dtype_code(q.dtype), _ptr(ws), 0 if ws is None else ws.numel(), _stream(q))
    _check(rc, "hg_prefix_attn_grouped_fwd")


def prefix_schedule(levels, n_q_rows: int, hq: int, n_sms: int = 148, allow_split=True):
    """Host-side work schedule (no GPU needed).  ``levels``: list of (n_groups, k_len, max_k_len) with max_k_len > 0 for
    ragged levels.  ``allow_split``: False / 0 whole units only, True / 1 what a launch does (stream-K where it pays),
    2 always stream-K.  Returns (n_ctas, list of piece tuples (cta, unit, level, head, group, tile, b_lo, b_hi, split, slot))."""
    arr = (PrefixLevel * len(levels))()
    for a, (ng, kl, mk) in zip(arr, levels):
        a.n_groups, a.k_len, a.max_k_len = ng, kl, mk
    lib = load()
    n_ctas = c_int32(0)
    n = lib.hg_prefix_schedule(arr, len(levels), n_q_rows, hq, n_sms, int(allow_split), None, 0, ctypes.byref(n_ctas))
    _check(min(n, 0), "hg_prefix_schedule")
    buf = (c_int32 * (10 * max(1, n)))()
    n2 = lib.hg_prefix_schedule(arr, len(levels), n_q_rows, hq, n_sms, int(allow_split), buf, n, ctypes.byref(n_ctas))
    _check(min(n2, 0), "hg_prefix_schedule")
    return int(n_ctas.value), [tuple(buf[i * 10 : i * 10 + 10]) for i in range(n2)]


def causal_attn_fwd(q, k, v, out, lse, b, sq, sk, hq, hkv, d, q_stride_row, kv_stride_row, sm_scale) -> None:
    ensure_init(q.device)
    with torch.cuda.device(q.device):
        rc = load().hg_causal_attn_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(lse), b, sq, sk, hq, hkv, d, q_stride_row,
                                       kv_stride_row, float(sm_scale), dtype_code(q.dtype), _stream(q))
    _check(rc, "hg_causal_attn_fwd")


def kv_append(k_new, v_new, positions, k_cache, v_cache) -> None:
    ensure_init(k_new.device)
    b, nq, hkv, d = k_new.shape
    lk = k_cache.shape[1]
    if positions.dtype == torch.int64:
        pi64 = 1
    elif positions.dtype == torch.int32:
        pi64 = 0
    else:
        raise ValueError(f"positions must be int32 or int64, got {positions.dtype}")
    with torch.cuda.device(k_new.device):
        rc = load().hg_kv_append(_ptr(k_new), _ptr(v_new), _ptr(positions), pi64, _ptr(k_cache), _ptr(v_cache),
                                 b, nq, lk, hkv, d, dtype_code(k_new.dtype), _stream(k_new))
    _check(rc, "hg_kv_append")


def decode_attn_fused(q, k_new, v_new, positions, k_cache, v_cache, out, lse, partial_outs, partial_lses, sm_scale) -> None:
    ensure_init(q.device)
    b, nq, hq, d = q.shape
    assert nq == 1
    hkv, lk = k_cache.shape[-2], k_cache.shape[1]
    if positions.dtype == torch.int64:
        pi64 = 1
    elif positions.dtype == torch.int32:
        pi64 = 0
    else:
        raise ValueError(f"positions must be int32 or int64, got {positions.dtype}")
    with torch.cuda.device(q.device):
        rc = load().hg_decode_attn_fused(
            _ptr(q), _ptr(k_new), _ptr(v_new), _ptr(positions), pi64, _ptr(k_cache), _ptr(v_cache), _ptr(out), _ptr(lse),
            b, lk, hq, hkv, d, q.stride(0), q.stride(2), k_cache.stride(0), k_cache.stride(1), k_cache.stride(2),
            _ptr_table(partial_outs), _ptr_table(partial_lses), len(partial_outs), float(sm_scale), dtype_code(q.dtype), _stream(q))
    _check(rc, "hg_decode_attn_fused")


def allreduce_multimem(mc_ptr: int, out_ptr: int, flags_dev: int, rank: int, world: int, nbytes: int, dtype: torch.dtype, n_blocks: int,
                       device) -> None:
    ensure_init(device)
    with torch.cuda.device(device):
        rc = load().hg_allreduce_multimem(c_void_p(mc_ptr), c_void_p(out_ptr), c_void_p(flags_dev), rank, world, nbytes, dtype_code(dtype), n_blocks,
                                          c_void_p(torch.cuda.current_stream(device).cuda_stream))
    _check(rc, "hg_allreduce_multimem")


def oproj_allreduce_flag_words(m: int, n: int, world: int) -> int:
    return int(load().hg_oproj_allreduce_flag_words(m, n, world))


def oproj_allreduce_plan(m: int, n: int, world: int, rank: int, n_ctas: int = 0, cover=None) -> dict:
    """Host-side view of the fused o_proj + all-reduce launch (hg_oproj_allreduce_plan; no device work).  ``cover``: a zeroed
    int32 numpy array of m * n / 8 entries that receives how often ``rank`` reduces each 16-byte vector of the output."""
    geo = (c_int * 8)()
    rc = load().hg_oproj_allreduce_plan(m, n, world, rank, n_ctas, ctypes.cast(geo, c_void_p),
                                        c_void_p(cover.ctypes.data) if cover is not None else None)
    _check(rc, "hg_oproj_allreduce_plan")
    keys = ("bn", "u", "reduce_warps", "ctas", "tiles", "tiles_owned", "slices_owned", "flag_words")
    return dict(zip(keys, list(geo)))


def oproj_allreduce_fwd(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, out_mc: int = 0, flags_dev: int = 0, flag_words: int = 0,
                        rank: int = 0, world: int = 1, n_ctas: int = 0) -> None:
    """out[m, n] = sum over ranks of x[m, k] @ w[n, k]^T (hg_oproj_allreduce_fwd; world == 1: the tcgen05 GEMM alone)."""
    ensure_init(x.device)
    if x.dim() != 2 or w.dim() != 2 or out.dim() != 2 or x.shape[1] != w.shape[1] or out.shape != (x.shape[0], w.shape[0]):
        raise ValueError(f"oproj_allreduce: x {tuple(x.shape)}, w {tuple(w.shape)}, out {tuple(out.shape)}")
    if not (x.dtype == w.dtype == out.dtype) or x.stride(1) != 1 or w.stride(1) != 1 or not out.is_contiguous():
        raise ValueError("oproj_allreduce: one 16-bit dtype, unit inner strides, contiguous out")
    with torch.cuda.device(x.device):
        rc = load().hg_oproj_allreduce_fwd(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(out), c_void_p(out_mc), c_void_p(flags_dev),
                                           flag_words, rank, world, x.shape[0], w.shape[0], x.shape[1], dtype_code(x.dtype), n_ctas, _stream(x))
    _check(rc, "hg_oproj_allreduce_fwd")


def rope_qk(q, k, q_out, k_out, cos_table, sin_table, positions, rows, hq, hkv, d, q_stride_row, k_stride_row,
            q_out_stride_row, k_out_stride_row) -> None:
    ensure_init(q.device)
    if positions.dtype == torch.int64:
        pi64 = 1
    elif positions.dtype == torch.int32:
        pi64 = 0
    else:
        raise ValueError(f"positions must be int32 or int64, got {positions.dtype}")
    with torch.cuda.device(q.device):
        rc = load().hg_rope_qk(_ptr(q), _ptr(k), _ptr(q_out), _ptr(k_out), _ptr(cos_table), _ptr(sin_table), _ptr(positions), pi64,
                               rows, hq, hkv, d, q_stride_row, k_stride_row, q_out_stride_row, k_out_stride_row,
                               cos_table.shape[0], dtype_code(q.dtype), _stream(q))
    _check(rc, "hg_rope_qk")
