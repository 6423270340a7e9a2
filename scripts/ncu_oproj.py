"""Development aid (1 GPU, run under ncu): a few launches of the o_proj tcgen05 GEMM alone (world == 1) at the shapes of a
decode step, for an `ncu --set full` capture of `oproj_allreduce_sm100_kernel`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200 import _lib  # noqa: E402

dev, dt = torch.device("cuda"), torch.bfloat16
for m, n, k in [(1024, 4096, 512), (1024, 4096, 4096)]:
    x = torch.randn(m, k, device=dev).to(dt)
    w = (torch.randn(n, k, device=dev) / k**0.5).to(dt)
    o = torch.empty(m, n, device=dev, dtype=dt)
    for _ in range(3):
        _lib.oproj_allreduce_fwd(x, w, o)
    torch.cuda.synchronize()
