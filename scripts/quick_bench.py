"""Kernel-level timings at BASELINE.json configs[1] (B=1024, prefix 2048, 32 heads, d=128, bf16):
our prefix / suffix+combine / whole operator vs the library kernels installed in the image
(flash-attn 2.8 FA2), CUDA events, L2 flushed between iterations (hydragen/benchmark_utils.py:82-137,
scripts/microbenchmark.py:24-47 of the reference).  Development tool, not the graded bench."""

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters=30, warmup=5, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return {"median_us": ts[len(ts) // 2], "min_us": ts[0], "max_us": ts[-1]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=1024)
    ap.add_argument("--ls", type=int, default=2048)
    ap.add_argument("--lu", type=int, default=1)
    ap.add_argument("--maxlu", type=int, default=16)
    ap.add_argument("--h", type=int, default=32)
    ap.add_argument("--hkv", type=int, default=32)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--no-lib", action="store_true")
    a = ap.parse_args()
    from hydragen_b200.attention import hydragen_attention_nopad
    from hydragen_b200.flash import flash_attention_seqlen, prefix_attention_grouped

    dev = "cuda:0"
    torch.manual_seed(0)
    dt = torch.bfloat16
    q = torch.randn(a.b, 1, a.h, a.d, device=dev, dtype=dt)
    k = torch.randn(a.b, a.maxlu, a.hkv, a.d, device=dev, dtype=dt)
    v = torch.randn_like(k)
    sk = torch.randn(1, a.ls, a.hkv, a.d, device=dev, dtype=dt)
    sv = torch.randn_like(sk)
    sl = torch.full((a.b,), a.lu, device=dev, dtype=torch.int64)
    flush = torch.empty(256 << 20, dtype=torch.int8, device=dev)
    res = {"config": vars(a)}
    flops = 4.0 * a.b * a.h * a.ls * a.d
    r = timed(lambda: prefix_attention_grouped(q, sk, sv, n_groups=1), flush=flush)
    r["tflops_median"] = flops / r["median_us"] / 1e6
    res["prefix_tcgen05"] = r
    os.environ["HYDRAGEN_B200_PREFIX_BACKEND"] = "rowwise"
    res["prefix_rowwise"] = timed(lambda: prefix_attention_grouped(q, sk, sv, n_groups=1), iters=5, warmup=1, flush=flush)
    os.environ["HYDRAGEN_B200_PREFIX_BACKEND"] = "auto"
    r = timed(lambda: flash_attention_seqlen(q, k, v, sl), flush=flush)
    r["gbs_median"] = (2.0 * a.b * a.lu * a.hkv * a.d * 2 + 2.0 * a.b * a.h * a.d * 2) / r["median_us"] / 1e3
    res["suffix_rowwise"] = r
    res["operator"] = timed(lambda: hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl), flush=flush)
    g = torch.cuda.CUDAGraph()
    hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl)
    res["operator_graph"] = timed(g.replay, flush=flush)
    if not a.no_lib:
        try:
            from flash_attn import flash_attn_func, flash_attn_with_kvcache

            r = timed(lambda: flash_attn_func(q.view(1, a.b, a.h, a.d), sk, sv, softmax_scale=a.d**-0.5, causal=False), flush=flush)
            r["tflops_median"] = flops / r["median_us"] / 1e6
            res["lib_fa2_prefix"] = r
            sl32 = sl.to(torch.int32)
            res["lib_fa2_suffix_kvcache"] = timed(lambda: flash_attn_with_kvcache(q, k, v, cache_seqlens=sl32, softmax_scale=a.d**-0.5), flush=flush)
        except Exception as ex:  # library baseline only
            res["lib_error"] = repr(ex)[:300]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
