"""Development aid (1 GPU; library built with HG_EXTRA_NVCC_FLAGS=-DHG_PREFIX_TRACE and the same variable set at run time):
%globaltimer stamps of the persistent prefix kernel's stages per CTA, for the last launch of a graph of back-to-back
launches at cfg#2 (or TP_B / TP_L / TP_H): where a CTA's time goes -- set-up, wait for the previous grid, first scores,
main loop, epilogue, merge, exit."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200 import _lib  # noqa: E402
from hydragen_b200.flash import prefix_attention_grouped  # noqa: E402

B, Lp, H, D, NL = int(os.environ.get("TP_B", "1024")), int(os.environ.get("TP_L", "2048")), int(os.environ.get("TP_H", "32")), 128, 8
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
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
lib = _lib.load()
n_ctas, pieces = _lib.prefix_schedule([(1, Lp, 0)], B, H, allow_split=os.environ.get("HYDRAGEN_B200_PREFIX_SPLIT", "1") != "0")
buf = (ctypes.c_longlong * (160 * 16))()
lib.hg_debug_prefix_trace(buf, 160 * 16)
rows = [[buf[c * 16 + s] for s in range(11)] for c in range(n_ctas)]
t0 = min(r[0] for r in rows)
names = ["entry", "set-up done", "dep-wait passed", "p1 first scores", "p1 loop done", "p1 epilogue done", "pN first scores", "pN loop done",
         "pN epilogue done", "merge done", "CTA done"]
print(f"B={B} L={Lp} H={H} split={os.environ.get('HYDRAGEN_B200_PREFIX_SPLIT', '1')}: {n_ctas} CTAs, {len(pieces)} pieces; us after the first CTA's entry: min / median / max over CTAs")
for s, nm in enumerate(names):
    vals = sorted((r[s] - t0) / 1e3 for r in rows if r[s] >= t0)
    if vals:
        print(f"  {nm:18s} {vals[0]:8.2f} {vals[len(vals) // 2]:8.2f} {vals[-1]:8.2f}   (n={len(vals)})")
