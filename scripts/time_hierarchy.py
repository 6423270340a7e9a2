"""BASELINE.json configs[3] at the operator level: the two-level hierarchy 1 x 1024 -> 32 x 64 -> 32 completions each
(B = 1024), Llama-2-7B head configuration (32 heads, d = 128, bf16), one decode step of one layer:
    ONE persistent prefix launch over both shared levels  +  ONE fused append / suffix / 3-way combine launch
(the reference: two flash-attn launches + two LSE transposes + cast + split-K + reduce + eager-torch 3-way combine,
hydragen/attention.py:250-352).  Graph-timed over `NL` layers' worth of distinct tensors, L2 flushed before every timed
replay.  Also timed: the prefix branch level by level (one launch per level), for the gain of the grouped launch.
`run()` returns a dict; as a script it prints it as JSON."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(B=1024, l1=1024, n2=32, l2=64, H=32, D=128, suffix=16, NL=4, iters=12, peaks=None):
    from hydragen_b200.attention import hydragen_attention_decode
    from hydragen_b200.flash import prefix_attention_grouped, prefix_attention_levels

    dev, dt = torch.device("cuda"), torch.bfloat16
    mk = lambda *s: torch.randn(*s, device=dev, dtype=dt)
    maxlu = (suffix + 15) // 16 * 16
    q = [mk(B, 1, H, D) for _ in range(NL)]
    kn, vn = [mk(B, 1, H, D) for _ in range(NL)], [mk(B, 1, H, D) for _ in range(NL)]
    kc, vc = [mk(B, maxlu, H, D) for _ in range(NL)], [mk(B, maxlu, H, D) for _ in range(NL)]
    s1k, s1v = [mk(1, l1, H, D) for _ in range(NL)], [mk(1, l1, H, D) for _ in range(NL)]
    s2k, s2v = [mk(n2, l2, H, D) for _ in range(NL)], [mk(n2, l2, H, D) for _ in range(NL)]
    pos = torch.full((B,), suffix - 1, device=dev, dtype=torch.int64)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / NL)
        ts.sort()
        return ts[len(ts) // 2]

    t_op = timed(lambda: [hydragen_attention_decode(q[i], kn[i], vn[i], pos, kc[i], vc[i], [s1k[i], s2k[i]], [s1v[i], s2v[i]]) for i in range(NL)])
    t_grouped = timed(lambda: [prefix_attention_levels(q[i], [s1k[i], s2k[i]], [s1v[i], s2v[i]], [1, n2], [None, None], [None, None]) for i in range(NL)])
    t_l1 = timed(lambda: [prefix_attention_grouped(q[i], s1k[i], s1v[i], n_groups=1) for i in range(NL)])
    t_l2 = timed(lambda: [prefix_attention_grouped(q[i], s2k[i], s2v[i], n_groups=n2) for i in range(NL)])
    flops = 4.0 * B * H * D * (l1 + l2)
    kv_bytes = 2.0 * (l1 + n2 * l2) * H * D * 2 + 2.0 * B * H * D * 2 * 2  # both levels' K/V once, q read once, two partial outs written
    res = {"config": f"B={B}: 1 x {l1} -> {n2} x {l2} -> {B // n2} completions each, {H} heads d={D}, bf16, suffix {suffix}",
           "operator_us": t_op, "prefix_grouped_launch_us": t_grouped, "prefix_level1_alone_us": t_l1, "prefix_level2_alone_us": t_l2,
           "prefix_level_by_level_us": t_l1 + t_l2, "prefix_flop": flops, "prefix_tflops": flops / t_grouped / 1e6,
           "prefix_min_bytes": kv_bytes, "prefix_gbs": kv_bytes / t_grouped / 1e3,
           "method": f"graph of {NL} launches on distinct tensors, L2 flushed (256 MiB) before every timed replay, median of {iters}"}
    if peaks:
        res["prefix_frac_of_tensor_peak"] = res["prefix_tflops"] / peaks.get("bf16_tflops", 1590.0)
        res["prefix_frac_of_hbm_peak"] = res["prefix_gbs"] / peaks.get("hbm_gbs", 6650.0)
    return res


if __name__ == "__main__":
    peaks = None
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    print(json.dumps(run(peaks=peaks)))
