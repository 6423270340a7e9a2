"""Development aid (1 GPU): the tcgen05 GEMM of csrc/oproj_allreduce.cu on its own (world == 1) against cuBLAS on the o_proj
shapes of a decode step, graph-timed over 8 distinct operand sets."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200 import _lib  # noqa: E402

dev = torch.device("cuda")
NL, dt = 8, torch.bfloat16


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * NL)


for m, n, k in [(1024, 4096, 512), (1024, 4096, 2048), (1024, 4096, 4096), (2048, 5120, 640), (2048, 5120, 5120)]:
    xs = [torch.randn(m, k, device=dev).to(dt) for _ in range(NL)]
    ws = [(torch.randn(n, k, device=dev) / k**0.5).to(dt) for _ in range(NL)]
    outs = [torch.empty(m, n, device=dev, dtype=dt) for _ in range(NL)]
    t_own = timed(lambda: [_lib.oproj_allreduce_fwd(x, w, o) for x, w, o in zip(xs, ws, outs)])
    t_lib = timed(lambda: [torch.matmul(x, w.t(), out=o) for x, w, o in zip(xs, ws, outs)])
    fl = 2.0 * m * n * k
    by = 2.0 * (m * k + n * k + m * n)
    print(f"[{m},{k}] x [{n},{k}]^T: own {t_own:.1f} us ({fl / t_own * 1e-6:.0f} TFLOP/s, {by / t_own * 1e-3:.0f} GB/s) | cuBLAS {t_lib:.1f} us", flush=True)
