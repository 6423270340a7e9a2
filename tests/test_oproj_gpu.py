"""The tcgen05 GEMM of csrc/oproj_allreduce.cu on one GPU (world == 1: no collective): out = x @ w.T against a torch fp32
reference of the same op (hydragen/llama.py:592-594, o_proj on the local heads).  The multi-rank form is covered by
tests/test_multigpu_gpu.py (torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(m, n, k, dtype, x_stride=None, n_ctas=0, seed=0):
    from hydragen_b200 import _lib

    g = torch.Generator(device="cuda").manual_seed(seed)
    xs = x_stride or k
    xbuf = torch.randn(m, xs, device="cuda", dtype=torch.float32, generator=g).to(dtype)
    x = xbuf[:, :k]
    w = (torch.randn(n, k, device="cuda", dtype=torch.float32, generator=g) / k**0.5).to(dtype)
    out = torch.full((m, n), float("nan"), device="cuda", dtype=dtype)
    _lib.oproj_allreduce_fwd(x, w, out, n_ctas=n_ctas)
    ref = x.float() @ w.float().t()
    torch.cuda.synchronize()
    # tolerance: one rounding of the fp32 accumulator to the 16-bit output type (2^-8 relative for bf16) + accumulation order
    tol = (1.0 / 128 if dtype == torch.bfloat16 else 1.0 / 1024) * ref.abs().max().item() + 1e-6
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= tol


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize(
    "m,n,k",
    [
        (1024, 4096, 512),   # cfg#2 at tp=8: one wave of 128 tiles
        (1024, 4096, 4096),  # single GPU: 64 k-blocks through the 3-deep ring
        (2048, 5120, 640),   # cfg#5 at tp=8: 320 tiles, both accumulators in use
        (1000, 520, 72),     # ragged in every dimension (zero-filled loads, clipped stores)
        (1, 8, 8),
        (130, 264, 136),
    ],
)
def test_gemm_matches_fp32_reference(m, n, k, dtype):
    _run(m, n, k, dtype)


def test_strided_rows_and_few_ctas():
    _run(512, 1024, 256, torch.bfloat16, x_stride=384)          # x = a column slice of a wider tensor
    _run(1024, 2048, 320, torch.bfloat16, n_ctas=5)              # 64 tiles on 5 CTAs: long tile loop, ring / accumulator phases wrap
    _run(640, 1280, 192, torch.float16, n_ctas=1)


def test_rejects_bad_arguments():
    from hydragen_b200 import _lib

    x = torch.zeros(16, 24, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(32, 24, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        _lib.oproj_allreduce_fwd(x, w, torch.zeros(16, 31, device="cuda", dtype=torch.bfloat16))
    with pytest.raises(RuntimeError):  # fp32 is not a tensor-core input type here
        _lib.oproj_allreduce_fwd(x.float(), w.float(), torch.zeros(16, 32, device="cuda"))
    with pytest.raises(RuntimeError):  # k not a multiple of 8
        _lib.oproj_allreduce_fwd(x[:, :20], w[:, :20], torch.zeros(16, 32, device="cuda", dtype=torch.bfloat16))


@pytest.mark.skipif(os.environ.get("HYDRAGEN_B200_OPROJ_BN") is not None, reason="already the narrow-tile run")
def test_narrow_tiles():
    """The 128-column tiling (what multi-rank launches of one-wave products use; a per-process switch) on the same cases."""
    env = dict(os.environ, HYDRAGEN_B200_OPROJ_BN="128")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", os.path.abspath(__file__), "-k",
                        "gemm_matches or strided"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0 and " passed" in r.stdout, (r.stdout + r.stderr)[-3000:]
