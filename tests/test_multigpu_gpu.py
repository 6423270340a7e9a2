"""Multi-rank GPU tests (``-m gpu``; each case is skipped when the box has fewer GPUs than it needs): one process per GPU
under torchrun, NCCL rendezvous on 127.0.0.1.

* the library's NVLS all-reduce kernel (csrc/allreduce.cu: multimem.ld_reduce / multimem.st, epoch-flag barriers) gives
  NCCL's result -- the collective of the reference's tensor-parallel path, ``funcol.all_reduce`` in hydragen/tp.py:108-112;
  bf16 sums in a different order: |diff| <= 2e-2 * max |value| (one bf16 ulp of the largest sum);
* head-axis tensor-parallel ``generate`` (hydragen/tp.py:30-124) on the CUDA kernels, eager and CUDA-graph decode,
  reproduces the unsharded model's logits (teacher-forced): |diff| <= 5e-2 * max |logit| in bf16 -- with o_proj / down_proj
  and their all-reduce as the fused launch (csrc/oproj_allreduce.cu);
* that fused launch on its own (tcgen05 GEMM tiles reduced by the switch as they are produced; hydragen/llama.py:592-594 +
  hydragen/tp.py:108-112) == per-rank matmul summed with NCCL, first call / repeated on the same buffers / graph replay /
  20 checked back-to-back rounds (a stale read of a partial would leak the previous sum): |diff| <= max |sum| / 64 in bf16.
"""

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(world, script, timeout=420, env=None):
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, **(env or {})), cwd=ROOT)
    return r.returncode, r.stdout + r.stderr


def _need(world):
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip(f"needs {world} GPUs, this box has {n}")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nvls_allreduce_matches_nccl(world):
    _need(world)
    rc, out = _torchrun(world, "check_allreduce.py")
    if "multicast support: False" in out:
        pytest.skip("no NVLink multicast on this box")
    assert rc == 0 and "parity ok" in out, out[-3000:]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_tp_generate_matches_unsharded_model(world):
    _need(world)
    rc, out = _torchrun(world, "check_tp_gpu.py")
    assert rc == 0 and "tp parity ok" in out, out[-3000:]
    assert f"tp={world} graph=True" in out


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_oproj_allreduce_matches_matmul_plus_nccl(world):
    _need(world)
    rc, out = _torchrun(world, "check_oproj_allreduce.py", env={"OPROJ_STRESS": "20", "OPROJ_FUSED_ONLY": "1"})
    assert rc == 0 and "oproj_allreduce parity ok" in out, out[-3000:]
