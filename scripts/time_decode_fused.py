"""Development aid (1 GPU): the fused decode launch (kv append + suffix + combine) of cfg#2 on its own and behind the prefix
launch, graph-timed over 32 layers' distinct tensors.  (Written for an A/B that is recorded in DESIGN.md 4.3 and not kept: CTAs
beyond the first resident wave fetching the prefix partial up front, behind an early griddepcontrol.wait -- 14.4 vs 14.1 us.)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200.flash import decode_attention_fused, prefix_attention_partials  # noqa: E402

dev, dt = torch.device("cuda"), torch.bfloat16
B, H, D, L, LP, LK = 1024, 32, 128, 32, 2048, 128
mk = lambda *s: torch.randn(*s, device=dev, dtype=dt)
qs, kn, vn = [mk(B, 1, H, D) for _ in range(L)], [mk(B, 1, H, D) for _ in range(L)], [mk(B, 1, H, D) for _ in range(L)]
sk, sv = [mk(1, LP, H, D) for _ in range(L)], [mk(1, LP, H, D) for _ in range(L)]
uniq = torch.randn(L, 2, B, LK, H, D, device=dev, dtype=dt)
pre = [prefix_attention_partials(qs[i], sk[i], sv[i], 1, max_splits=1) for i in range(L)]


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * L)


for suffix, rows in ((1, LK), (1, 32), (1, 1), (2, LK), (2, 32), (16, LK), (16, 32), (16, 16), (64, LK)):
    pos = torch.full((B, 1), suffix - 1, device=dev, dtype=torch.int64)
    # `rows`: how many rows of the 128-row caches the call is shown (a view; <= 32 selects the short-cache variant of the kernel)
    kc = [uniq[i, 0][:, :rows] for i in range(L)]
    vc = [uniq[i, 1][:, :rows] for i in range(L)]

    def only_suffix():
        for i in range(L):
            decode_attention_fused(qs[i], kn[i], vn[i], pos, kc[i], vc[i], pre[i][0], pre[i][1])

    def step():
        for i in range(L):
            o, l = prefix_attention_partials(qs[i], sk[i], sv[i], 1, max_splits=1)
            decode_attention_fused(qs[i], kn[i], vn[i], pos, kc[i], vc[i], o, l)

    print(f"suffix {suffix}, cache view of {rows} rows: {timed(only_suffix):.2f} us alone, {timed(step):.2f} us per layer with the prefix launch", flush=True)
