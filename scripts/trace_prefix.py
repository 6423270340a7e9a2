"""Development aid: per-block clock64 timeline of one prefix-kernel CTA (needs a build with
HG_EXTRA_NVCC_FLAGS=-DHG_PREFIX_TRACE, which goes to hydragen_b200/_C_dev).  Run under gpurun:
    HG_EXTRA_NVCC_FLAGS=-DHG_PREFIX_TRACE python scripts/trace_prefix.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200 import _lib  # noqa: E402
from hydragen_b200.flash import prefix_attention_grouped  # noqa: E402

lib = _lib.load()
B, L, H, D = int(os.environ.get("TRACE_B", "1024")), 2048, 32, 128
q = torch.randn(B, 1, H, D, device="cuda", dtype=torch.bfloat16)
k = torch.randn(1, L, H, D, device="cuda", dtype=torch.bfloat16)
v = torch.randn(1, L, H, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    prefix_attention_grouped(q, k, v, n_groups=1)
torch.cuda.synchronize()
n = 3 * 64 * 8
buf = (ctypes.c_longlong * n)()
lib.hg_debug_read_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = lib.hg_debug_read_trace(buf, n)
assert rc == 0, rc
tr = [[[buf[(r * 64 + j) * 8 + s] for s in range(8)] for j in range(64)] for r in range(3)]
nb = (L + 63) // 64
t0 = min(x for x in tr[1][0][:2] if x > 0)
print("MMA warp A: j | wait_kv wait_p issue+commit | iter start")
for j in range(min(nb, 32)):
    m = tr[0][j]
    print(f"{j:3d} | {m[1]-m[0]:6d} {m[2]-m[1]:6d} {m[3]-m[2]:6d} | {m[0]-t0:7d}")
for r, nm in ((1, "A"), (2, "B")):
    print(f"softmax {nm}: j | exp groups 0-5 | wait S(j+1)+LDTM issue | exp 6-7, max next, pack | st+arrive | total | start")
    for j in range(min(nb, 32)):
        s = tr[r][j]
        print(f"{j:3d} | {s[2]-s[1]:6d} {s[3]-s[2]:6d} {s[4]-s[3]:6d} {s[5]-s[4]:6d} | {s[5]-s[1]:6d} | {s[1]-t0:7d}")
