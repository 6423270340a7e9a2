"""Development aid: graph-timed prefix kernel at cfg#2 (one launch per layer on distinct tensors), us per launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200.flash import prefix_attention_grouped  # noqa: E402

B, Lp, H, D, NL = int(os.environ.get("TP_B", "1024")), int(os.environ.get("TP_L", "2048")), int(os.environ.get("TP_H", "32")), 128, 16
q = [torch.randn(B, 1, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(NL)]
k = [torch.randn(1, Lp, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(NL)]
v = [torch.randn(1, Lp, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(NL)]


def run():
    for i in range(NL):
        prefix_attention_grouped(q[i], k[i], v[i], n_groups=1)


run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (20 * NL)
print(f"{os.environ.get('HG_EXTRA_NVCC_FLAGS', '-')} softmax={os.environ.get('HYDRAGEN_B200_PREFIX_SOFTMAX', 'default')} B={B} L={Lp}: prefix {us:.2f} us/launch, {4.0 * B * H * Lp * D / us / 1e6:.0f} TFLOP/s")
