"""Development aid: per-CTA main-loop stamps of the ROUND-1 prefix kernel (instrumented build in
hydragen_b200/_C_r01trace, made from commit 7e4cab3 plus four %globaltimer / clock64 stamps), printed like
scripts/trace_prefix.py prints them for the current kernel -- same box, same call."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "hydragen_b200", "_C_r01trace", "libhydragen_b200.so")
if not os.path.exists(path):
    print("instrumented round-1 library not present: skipped")
    sys.exit(0)
lib = ctypes.CDLL(path)
B, Lp, H, D, NL = 1024, 2048, 32, 128, 8
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
        assert lib.hg_prefix_attn_fwd(q[i].data_ptr(), k[i].data_ptr(), v[i].data_ptr(), o[i].data_ptr(), l[i].data_ptr(), 1, B, Lp, Lp, None, Lp,
                                      H, H, D, H * D, H * D, D**-0.5, 1, st) == 0


run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    run()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (160 * 16))()
lib.hg_debug_prefix_trace(buf, 160 * 16)
rows = [[buf[c * 16 + s] for s in range(14)] for c in range(128)]
t0 = min(r[2] for r in rows)
for s, nm in [(2, "dep-wait passed"), (3, "first scores"), (4, "loop done"), (10, "CTA done")]:
    vals = sorted((r[s] - t0) / 1e3 for r in rows)
    print(f"  {nm:18s} {vals[0]:8.2f} {vals[len(vals) // 2]:8.2f} {vals[-1]:8.2f}")
print("  cycles per key block, by SM id (smid:cycles/block):")
by_sm = sorted((int(r[11]), (r[13] - r[12]) / 32.0) for r in rows)
line = []
for sm, cpb in by_sm:
    line.append(f"{sm}:{cpb:.0f}")
    if len(line) == 16:
        print("    " + " ".join(line))
        line = []
if line:
    print("    " + " ".join(line))
