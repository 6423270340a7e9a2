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
rows = [[buf[c * 16 + s] for s in range(16)] for c in range(n_ctas)]
t0 = min(r[0] for r in rows)
names = ["entry", "set-up done", "dep-wait passed", "p1 first scores", "p1 loop done", "p1 epilogue done", "pN first scores", "pN loop done",
         "pN epilogue done", "merge done", "CTA done"]
print(f"B={B} L={Lp} H={H} split={os.environ.get('HYDRAGEN_B200_PREFIX_SPLIT', '1')}: {n_ctas} CTAs, {len(pieces)} pieces; us after the first CTA's entry: min / median / max over CTAs")
for s, nm in enumerate(names):
    vals = sorted((r[s] - t0) / 1e3 for r in rows if r[s] >= t0)
    if vals:
        print(f"  {nm:18s} {vals[0]:8.2f} {vals[len(vals) // 2]:8.2f} {vals[-1]:8.2f}   (n={len(vals)})")

# per-CTA view of the first piece: main-loop time, SM, cycles (SM clock = cycles / time)
det = sorted(((r[4] - r[3]) / 1e3, c, int(r[11]), r[13] - r[12], (r[3] - r[2]) / 1e3, (r[9] - r[2]) / 1e3, r[14], r[8], r[15]) for c, r in enumerate(rows) if r[4] > r[3] > 0)
print("  first piece, per CTA (fastest 6, slowest 12): loop us | cta | smid | loop cycles | MHz | dep-wait -> first scores us | dep-wait -> merge done us"
      " | whole-launch wait cycles: softmax A on S, MMA A on P, MMA A on K/V")
for lt, c, sm, cyc, st, tot, ws, wp, wkv in det[:6] + det[-12:]:
    print(f"    {lt:7.2f} {c:4d} {sm:4d} {cyc:8d} {cyc / max(lt, 1e-9):7.0f} {st:6.2f} {tot:7.2f} | {ws:7d} {wp:7d} {wkv:7d}")
first_blocks = {}
for p_ in pieces:
    first_blocks.setdefault(p_[0], p_[7] - p_[6])
print("  cycles per key block of the first piece, by SM id (smid:cycles/block):")
by_sm = sorted((int(r[11]), (r[13] - r[12]) / max(1, first_blocks.get(c, 1)), c) for c, r in enumerate(rows) if r[4] > r[3] > 0)
line = []
for sm, cpb, c in by_sm:
    line.append(f"{sm}:{cpb:.0f}")
    if len(line) == 16:
        print("    " + " ".join(line))
        line = []
if line:
    print("    " + " ".join(line))
