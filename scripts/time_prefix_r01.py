"""Development aid: the ROUND-1 prefix kernel (hydragen_b200/_C_r01/libhydragen_b200.so, built from commit 7e4cab3:
one CTA per (256-row tile, head), 128 CTAs at cfg#2) timed exactly like scripts/time_prefix.py, on the same box and in
the same call as the current kernel -- boxes differ by a few percent, so A/B numbers must come from one run."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "hydragen_b200", "_C_r01", "libhydragen_b200.so")
if not os.path.exists(path):
    print("round-1 library not present: skipped")
    sys.exit(0)
lib = ctypes.CDLL(path)
B, Lp, H, D, NL = int(os.environ.get("TP_B", "1024")), int(os.environ.get("TP_L", "2048")), int(os.environ.get("TP_H", "32")), 128, 16
q = [torch.randn(B, 1, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(NL)]
k = [torch.randn(1, Lp, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(NL)]
v = [torch.randn(1, Lp, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(NL)]
o = [torch.empty(B, 1, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(NL)]
l = [torch.empty(B, 1, H, device="cuda", dtype=torch.float32) for _ in range(NL)]
assert lib.hg_init(0) == 0
P, I, I64, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
lib.hg_prefix_attn_fwd.argtypes = [P, P, P, P, P, I, I, I64, I, P, I, I, I, I, I64, I64, F, I, P]


def run():
    st = torch.cuda.current_stream().cuda_stream
    for i in range(NL):
        rc = lib.hg_prefix_attn_fwd(q[i].data_ptr(), k[i].data_ptr(), v[i].data_ptr(), o[i].data_ptr(), l[i].data_ptr(), 1, B, Lp, Lp, None, Lp,
                                    H, H, D, H * D, H * D, D**-0.5, 1, st)
        assert rc == 0


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
print(f"round-1 kernel B={B} L={Lp} H={H}: prefix {us:.2f} us/launch, {4.0 * B * H * Lp * D / us / 1e6:.0f} TFLOP/s")
