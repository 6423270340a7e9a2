"""Target for ncu: a few launches of the grouped (persistent) prefix kernel at BASELINE.json configs[3] (two shared levels in one
launch) and of the fused decode launch at suffix 64."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200.flash import decode_attention_fused, prefix_attention_grouped, prefix_attention_levels  # noqa: E402

dev, dt = torch.device("cuda"), torch.bfloat16
mk = lambda *s: torch.randn(*s, device=dev, dtype=dt)
B, H, D = 1024, 32, 128
q = mk(B, 1, H, D)
s1k, s1v, s2k, s2v = mk(1, 1024, H, D), mk(1, 1024, H, D), mk(32, 64, H, D), mk(32, 64, H, D)
sk, sv = mk(1, 2048, H, D), mk(1, 2048, H, D)
kn, vn, kc, vc = mk(B, 1, H, D), mk(B, 1, H, D), mk(B, 128, H, D), mk(B, 128, H, D)
pos = torch.full((B,), 63, device=dev, dtype=torch.int64)
for _ in range(6):
    prefix_attention_levels(q, [s1k, s2k], [s1v, s2v], [1, 32], [None, None], [None, None])
    o, l = prefix_attention_grouped(q, sk, sv, n_groups=1)
    decode_attention_fused(q, kn, vn, pos, kc, vc, [o], [l])
torch.cuda.synchronize()
print("done")
