"""Attention-only microbenchmark: the re-creation of the reference's scripts/microbenchmark.py for this library.

Method = the reference's (hydragen/benchmark_utils.py:82-170, scripts/microbenchmark.py:24-47): random q / k / v,
the operator captured in a CUDA graph, per iteration [flush L2 by overwriting a buffer larger than it -> event ->
replay -> event -> synchronize], median over the timed iterations.  Sweeps (batch, prefix) pairs x unique suffix
lengths like docs/sweeps_from_paper.md:152-169; the reference's sweep DSL is kept ("a,b,c" | "start:end:step" |
"start:end:xK", hydragen/benchmark_utils.py:207-229).

Reported per config: the decode operator (prefix launch + fused append/suffix/combine launch), each kernel alone with
its roofline figure, and -- as the on-box point of comparison -- the flash-attn 2.8 (FA2, mma.sync) kernels the
reference would call for the same two branches (library code, baseline only).

    python scripts/microbenchmark.py --pairs 1024:2048 --suffix 1,64 --out profiles/microbenchmark.json
"""

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def split_range(spec: str):
    """hydragen/benchmark_utils.py:207-229: "a,b,c" | "start:end:step" | "start:end:xK" (geometric)."""
    if ":" not in spec:
        return [int(x) for x in spec.split(",") if x]
    start, end, step = spec.split(":")
    start, end = int(start), int(end)
    out = []
    if step.startswith("x"):
        k, v = float(step[1:]), float(start)
        while v <= end:
            out.append(int(round(v)))
            v *= k
    else:
        out = list(range(start, end + 1, int(step)))
    return out


def timed(fn, flush, iters=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def graphed(fn):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", default="512:1024,1024:2048,2048:4096,4096:8192", help="batch:prefix pairs")
    ap.add_argument("--suffix", default="1,16,64,128", help="unique tokens per sequence incl. the new one (sweep DSL)")
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--kv-heads", type=int, default=32)
    ap.add_argument("--head-dim", type=int, default=128)
    ap.add_argument("--no-lib", action="store_true", help="skip the flash-attn library baselines")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()

    from hydragen_b200.attention import hydragen_attention_decode
    from hydragen_b200.flash import decode_attention_fused, prefix_attention_grouped

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak, hbm_peak = float(peaks.get("bf16_tflops", 1590.0)), float(peaks.get("hbm_gbs", 6650.0))
    dev, dt = "cuda:0", torch.bfloat16
    H, HKV, D = a.heads, a.kv_heads, a.head_dim
    flush = torch.empty(256 << 20, dtype=torch.int8, device=dev)  # > the 126 MB L2
    rows = []
    for pair in a.pairs.split(","):
        B, Ls = (int(x) for x in pair.split(":"))
        for lu in split_range(a.suffix):
            torch.manual_seed(0)
            maxlu = (lu + 15) // 16 * 16
            q = torch.randn(B, 1, H, D, device=dev, dtype=dt)
            kn, vn = torch.randn(B, 1, HKV, D, device=dev, dtype=dt), torch.randn(B, 1, HKV, D, device=dev, dtype=dt)
            kc, vc = torch.randn(B, maxlu, HKV, D, device=dev, dtype=dt), torch.randn(B, maxlu, HKV, D, device=dev, dtype=dt)
            sk, sv = torch.randn(1, Ls, HKV, D, device=dev, dtype=dt), torch.randn(1, Ls, HKV, D, device=dev, dtype=dt)
            pos = torch.full((B,), lu - 1, device=dev, dtype=torch.int64)
            so, sl = prefix_attention_grouped(q, sk, sv, n_groups=1)
            t_op = timed(graphed(lambda: hydragen_attention_decode(q, kn, vn, pos, kc, vc, [sk], [sv])), flush)
            t_pre = timed(graphed(lambda: prefix_attention_grouped(q, sk, sv, n_groups=1)), flush)
            t_suf = timed(graphed(lambda: decode_attention_fused(q, kn, vn, pos, kc, vc, [so], [sl])), flush)
            flops = 4.0 * B * H * Ls * D
            nbytes = 4.0 * B * HKV * D * 2 + 2.0 * B * (lu - 1) * HKV * D * 2 + 3.0 * B * H * D * 2 + 2.0 * B * H * 4
            row = {"batch": B, "prefix": Ls, "suffix": lu, "operator_us": round(t_op, 2), "prefix_us": round(t_pre, 2),
                   "prefix_tflops": round(flops / t_pre / 1e6, 1), "prefix_frac_of_measured_peak": round(flops / t_pre / 1e6 / tf_peak, 3),
                   "suffix_side_us": round(t_suf, 2), "suffix_side_gbs": round(nbytes / t_suf / 1e3, 1),
                   "suffix_side_frac_of_measured_hbm": round(nbytes / t_suf / 1e3 / hbm_peak, 3)}
            if not a.no_lib:
                try:
                    from flash_attn import flash_attn_func, flash_attn_with_kvcache

                    sl32 = (pos + 1).to(torch.int32)
                    row["lib_fa2_prefix_us"] = round(timed(graphed(lambda: flash_attn_func(q.view(1, B, H, D), sk, sv, softmax_scale=D**-0.5)), flush, iters=10), 2)
                    row["lib_fa2_suffix_us"] = round(timed(graphed(lambda: flash_attn_with_kvcache(q, kc, vc, cache_seqlens=sl32, softmax_scale=D**-0.5)), flush, iters=10), 2)
                except Exception as ex:
                    row["lib_error"] = repr(ex)[:200]
            rows.append(row)
            print(json.dumps(row), flush=True)
            del q, kn, vn, kc, vc, sk, sv
    if a.out:
        json.dump({"method": "CUDA graph replay, L2 flushed (256 MiB) before every timed iteration, median of 30", "heads": H, "kv_heads": HKV,
                   "head_dim": D, "dtype": "bf16", "peaks": {"bf16_tflops": tf_peak, "hbm_gbs": hbm_peak}, "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
