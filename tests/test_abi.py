"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU
and exports every symbol include/hydragen_b200.h declares; the Python surface mirrors the
reference's names; the product path refuses to run without CUDA (no CPU fallback)."""

import ctypes
import inspect
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hydragen_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hg_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported(built_lib):
    names = _declared_symbols()
    assert {"hg_init", "hg_combine_lse", "hg_rowwise_attn_fwd", "hg_prefix_attn_fwd", "hg_kv_append", "hg_last_error"} <= set(names)
    lib = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hydragen_b200.h but not exported"
    lib.hg_abi_version.restype = ctypes.c_int
    assert lib.hg_abi_version() == 2


def test_binding_table_matches_header(built_lib):
    from hydragen_b200 import _lib

    assert sorted(_lib.SYMBOLS) == _declared_symbols()
    _lib.load()  # binds every symbol, checks the ABI version


def test_library_is_sm100a_with_tcgen05_and_tma(built_lib):
    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    r = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", built_lib], capture_output=True, text=True)
    assert r.returncode == 0
    assert "sm_100a" in r.stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):  # tcgen05.mma / ld / st, TMA load
        assert mnemonic in r.stdout, mnemonic
    assert "HMMA." not in r.stdout.replace("UTCHMMA", "")  # no legacy mma.sync path


def test_argument_validation_without_gpu(built_lib):
    """Invalid arguments are rejected before any CUDA call, so this runs on the CPU box."""
    from hydragen_b200 import _lib

    lib = _lib.load()
    rc = lib.hg_combine_lse(None, None, 0, None, None, 4, 64, 0, None)
    assert rc == -1 and b"n = 0" in lib.hg_last_error()
    rc = lib.hg_prefix_attn_fwd(None, None, None, None, None, 1, 1, 1, 1, None, 1, 8, 8, 128, 1024, 1024, 0.1, 1, None)
    assert rc == -4  # hg_init not called
    args = [None] * 4 + [0, None, 1, 0, None, None, 2, 1, 4, 8, 3, 128] + [0] * 6 + [None, None, 0, 0.1, 1, None]
    rc = lib.hg_rowwise_attn_fwd(*args)
    assert rc == -1 and b"multiple of hkv" in lib.hg_last_error()


def test_surface_mirrors_reference():
    import hydragen_b200.attention as A
    import hydragen_b200.flash as F

    sig = inspect.signature(A.hydragen_attention)
    assert list(sig.parameters) == ["q", "k", "v", "shared_ks", "shared_vs", "shared_cu_seq_lens", "shared_max_seq_lens", "use_varlens", "seq_lens"]
    assert list(inspect.signature(A.hydragen_attention_nopad).parameters) == ["q", "k", "v", "shared_ks", "shared_vs", "seq_len"]
    assert list(inspect.signature(A.combine_lse).parameters) == ["outs", "lses", "enable_triton"]
    assert list(inspect.signature(F.flash_attention).parameters) == ["q", "k", "v", "causal"]
    assert list(inspect.signature(F.flash_attention_varlen).parameters) == [
        "q", "k", "v", "cu_seqlens_q", "cu_seqlens_k", "max_seqlen_q", "max_seqlen_k", "causal"]
    assert list(inspect.signature(F.flash_attention_seqlen).parameters) == ["raw_q", "raw_k", "raw_v", "seq_len"]


def test_no_cpu_fallback():
    """CPU tensors must fail loudly: the product path never computes on the host."""
    import hydragen_b200.attention as A
    from hydragen_b200._lib import HydragenB200Error

    q = torch.randn(2, 1, 4, 64)
    k = torch.randn(2, 3, 4, 64)
    sk = torch.randn(1, 5, 4, 64)
    with pytest.raises(HydragenB200Error):
        A.hydragen_attention_nopad(q, k, k, [sk], [sk])
    with pytest.raises(HydragenB200Error):
        A.combine_lse([q, q], [torch.zeros(2, 1, 4), torch.zeros(2, 1, 4)])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "hydragen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"


def test_new_entry_points_validate_without_gpu(built_lib):
    """hg_rope_qk / hg_causal_attn_fwd reject bad arguments before any CUDA call."""
    from hydragen_b200 import _lib

    lib = _lib.load()
    # rope: head_dim must be a multiple of 16 for 16-bit dtypes
    rc = lib.hg_rope_qk(None, None, None, None, None, None, None, 1, 4, 2, 2, 24, 48, 48, 48, 48, 16, 1, None)
    assert rc == -2 and b"head_dim" in lib.hg_last_error()
    # rope: null tensors with rows > 0
    rc = lib.hg_rope_qk(None, None, None, None, None, None, None, 1, 4, 2, 2, 64, 128, 128, 128, 128, 16, 1, None)
    assert rc == -1 and b"null" in lib.hg_last_error()
    # rope: a row stride smaller than heads * head_dim
    buf = ctypes.c_void_p(4096)
    rc = lib.hg_rope_qk(buf, buf, buf, buf, buf, buf, buf, 1, 4, 2, 2, 64, 64, 128, 128, 128, 16, 1, None)
    assert rc == -1 and b"row stride" in lib.hg_last_error()
    # rope: zero rows is a no-op (no device needed)
    assert lib.hg_rope_qk(None, None, None, None, None, None, None, 0, 0, 2, 2, 64, 128, 128, 128, 128, 16, 1, None) == 0
    # causal: needs hg_init; after that sk < sq is unsupported -- without a GPU only the first is reachable
    rc = lib.hg_causal_attn_fwd(None, None, None, None, None, 1, 8, 4, 2, 2, 128, 256, 256, 0.1, 1, None)
    assert rc == -4


def test_oproj_allreduce_validates_without_gpu(built_lib):
    """hg_oproj_allreduce_fwd rejects bad arguments before any CUDA call; the flag-word count is a host-side formula."""
    from hydragen_b200 import _lib

    lib = _lib.load()
    # 128 header words + (128-row x 128-column tiles) x world
    assert lib.hg_oproj_allreduce_flag_words(1024, 4096, 8) == 128 + 8 * 32 * 8
    assert lib.hg_oproj_allreduce_flag_words(1, 8, 2) == 128 + 2
    assert lib.hg_oproj_allreduce_flag_words(-1, 8, 2) == 0
    buf = ctypes.c_void_p(4096)
    call = lambda **kw: lib.hg_oproj_allreduce_fwd(  # noqa: E731
        kw.get("x", buf), kw.get("xs", 512), kw.get("w", buf), kw.get("ws", 512), kw.get("out", buf), kw.get("mc", buf), kw.get("flags", buf),
        kw.get("words", 1 << 20), kw.get("rank", 0), kw.get("world", 2), kw.get("m", 1024), kw.get("n", 4096), kw.get("k", 512), kw.get("dtype", 1),
        kw.get("ctas", 0), None)
    assert call(dtype=2) == -2 and b"16-bit" in lib.hg_last_error()  # fp32 is not a tensor-core input type here
    assert call(rank=2) == -1 and b"rank" in lib.hg_last_error()
    assert call(k=508) == -2 and b"multiples of 8" in lib.hg_last_error()
    assert call(xs=256) == -2  # row stride smaller than k
    assert call(mc=None) == -1 and b"multicast" in lib.hg_last_error()
    assert call(words=64) == -1 and b"flag words" in lib.hg_last_error()
    assert call(m=0) == -1 and b"empty" in lib.hg_last_error()  # a rank may not skip a collective call
    assert call(x=None) == -1 and b"null" in lib.hg_last_error()
    assert call(world=1, m=0) == 0  # the GEMM alone on nothing: a no-op


def test_rope_and_causal_refuse_cpu_tensors():
    from hydragen_b200._lib import HydragenB200Error
    from hydragen_b200.flash import flash_attention
    from hydragen_b200.rope import apply_rotary_pos_emb

    q = torch.randn(2, 64, 4, 64, dtype=torch.bfloat16)
    cos = torch.zeros(128, 64, dtype=torch.bfloat16)
    pos = torch.zeros(2, 64, dtype=torch.long)
    with pytest.raises(HydragenB200Error):
        apply_rotary_pos_emb(q, q.clone(), cos, cos, pos)
    with pytest.raises(HydragenB200Error):
        flash_attention(q, q, q, causal=True)
    # argument errors are raised before the device is touched
    with pytest.raises(ValueError):
        apply_rotary_pos_emb(q, q.clone(), cos.float(), cos.float(), pos)
    with pytest.raises(NotImplementedError):
        apply_rotary_pos_emb(q, q.clone(), cos, cos, pos, unsqueeze_dim=1)


def test_row_view_helpers_accept_fused_projection_views():
    """Host logic of the shims: which strided [b, s, h, d] views can be handed to the kernels without a copy."""
    from hydragen_b200.flash import _rows_view
    from hydragen_b200.rope import _row_stride

    qkv = torch.zeros(6, 1, 3 * 4 * 64)
    q = qkv[..., : 4 * 64].unflatten(-1, (4, 64))          # decode: q rows inside a fused qkv buffer
    assert _rows_view(q) == 3 * 4 * 64 and _row_stride(q) == 3 * 4 * 64
    x = torch.zeros(2, 5, 4, 64)
    assert _rows_view(x) == 256 and _row_stride(x) == 256   # dense
    assert _rows_view(x[:, :3]) is None and _row_stride(x[:, :3]) is None      # (b, s) axes do not collapse
    assert _rows_view(x[:1, :3]) == 256 and _row_stride(x[:1, :3]) == 256      # single batch entry: fine
    assert _rows_view(x[:, :, ::2]) is None and _row_stride(x[:, :, ::2]) is None  # heads not dense
    assert _rows_view(x.transpose(1, 2)) is None
    cache = torch.zeros(3, 16, 2, 64)
    assert _rows_view(cache[:, :9]) is None                 # a cache prefix slice needs a copy ...
    assert _rows_view(cache[:1, :9]) == 128                 # ... unless it is one sequence


def test_header_is_plain_c_and_links(built_lib, tmp_path):
    """include/hydragen_b200.h compiles as C99 (no C++-isms, no torch types) and a C program linked against the
    library can call the entry points that need no GPU."""
    import shutil

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "hydragen_b200.h"\n'
        "int main(void) {\n"
        "  if (hg_abi_version() != HG_ABI_VERSION) return 1;\n"
        "  if (hg_sm_count() != 0) return 2;                       /* hg_init not called */\n"
        "  if (hg_combine_lse(0, 0, 0, 0, 0, 4, 64, HG_BF16, 0) != HG_ERR_INVALID_ARGUMENT) return 3;\n"
        "  if (strstr(hg_last_error(), \"n = 0\") == 0) return 4;\n"
        "  if (hg_prefix_workspace_bytes() < 4096) return 5;\n"
        "  if (hg_prefix_suggest_splits(1, 1024, 4, 2048, HG_MAX_COMBINE) < 1) return 8;\n"
        "  { hg_prefix_level lv; int32_t n_ctas = 0; memset(&lv, 0, sizeof lv); lv.n_groups = 1; lv.k_len = 2048;\n"
        "    if (hg_prefix_schedule(&lv, 1, 1024, 32, 148, 2, 0, 0, &n_ctas) < 128 || n_ctas != 148) return 6;\n"
        "    if (hg_prefix_schedule(&lv, 1, 1024, 32, 148, 0, 0, 0, &n_ctas) != 128 || n_ctas != 128) return 7; }\n"
        '  printf("abi %d ok\\n", hg_abi_version());\n'
        "  return 0;\n}\n")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(built_lib)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lhydragen_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "abi 2 ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
