#!/usr/bin/env python
"""bench.py -- decode throughput of the Hydragen shared-prefix attention hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the microbenchmark shape, for every layer of Llama-2-7B):
B = 1024 sequences decoding one token each against ONE shared prefix of 2048 tokens, every
sequence owning `--suffix-len` tokens of its own KV (default 1), 32 query / 32 kv heads, d = 128,
bf16.  One STEP = the attention hot path of one whole-model decode step: for each of the 32 layers
(each with its OWN caches, so 1.5 GB+ of distinct inputs stream through the 126 MB L2 per step)
    prefix attention (tcgen05)  ->  ONE launch: KV append of the new token + suffix attention + combine
i.e. what the reference's decode branch does with update_per_completion_kvs + hydragen_attention
(hydragen/llama.py:564-587), through ``hydragen_attention_decode``.
The step is captured in a CUDA graph (the reference replays graphs too: llama.py:781-866,
benchmark_utils.py:140-170).  `value` = B / step time = decode tokens/s of the attention path
(the projections / MLP / sampling around it are out of scope: SURVEY.md section 8).

Multi-GPU = the reference's head-axis tensor parallelism (hydragen/tp.py): each rank runs the same
step on Hq/N local heads, then per layer ONE NCCL all-reduce of the [B, hidden] bf16 tensor that
the row-parallel o_proj would produce.  Total work is fixed -> "scaling": "strong".

`e2e`: the same step, but every layer's q / k_new / v_new come from pinned HOST memory and the
attention output is read back to the host inside the timed region.
`--impl reference`: the CPU restatement of the reference's algorithm (oracle/, torch fp32 on all
host cores; the reference itself has no CPU attention path and cannot run here: DESIGN.md).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "decode tokens/sec (Llama-2-7B, 2K shared prefix, bs=1024)"
UNIT = "tokens/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--prefix-len", type=int, default=2048)
    ap.add_argument("--suffix-len", type=int, default=1, help="valid unique tokens per sequence (incl. the new one)")
    ap.add_argument("--max-unique-len", type=int, default=16, help="unique cache length (setup_caches rounds to 16)")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--kv-heads", type=int, default=32)
    ap.add_argument("--head-dim", type=int, default=128)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-model", action="store_true",
                    help="also run scripts/full_model_decode.py (random-init Llama-2-7B generate, configs[2]) and attach its numbers")
    return ap.parse_args()


def workload_name(a):
    return (f"microbenchmark cfg#2 x {a.layers} layers: B={a.batch}, shared prefix {a.prefix_len}, suffix {a.suffix_len}, "
            f"{a.heads}q/{a.kv_heads}kv heads d={a.head_dim}, bf16")


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle = restatement of the reference's algorithm; test/bench infrastructure only)
# ------------------------------------------------------------------------------------------------


def cpu_layer_inputs(a, torch, heads=None, kv_heads=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    h, hk = heads or a.heads, kv_heads or a.kv_heads
    mk = lambda *s: torch.randn(*s, generator=g, dtype=torch.float32)
    q = mk(a.batch, 1, h, a.head_dim)
    k = mk(a.batch, a.max_unique_len, hk, a.head_dim)
    v = mk(a.batch, a.max_unique_len, hk, a.head_dim)
    sk = mk(1, a.prefix_len, hk, a.head_dim)
    sv = mk(1, a.prefix_len, hk, a.head_dim)
    sl = torch.full((a.batch,), a.suffix_len, dtype=torch.int64)
    return q, k, v, sk, sv, sl


def time_cpu_layer(a, torch, budget_s):
    """Times ONE layer's hydragen_attention (fp32, all host threads) repeatedly for ~budget_s; a
    decode step is `layers` such calls, so tokens/s = B / (layers * t_layer)."""
    from oracle import hydragen_oracle as O

    q, k, v, sk, sv, sl = cpu_layer_inputs(a, torch)
    run = lambda: O.hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl, compute_dtype=torch.float32)
    run()  # warm-up
    ts = []
    t_end = time.perf_counter() + budget_s
    while len(ts) < 3 or (time.perf_counter() < t_end and len(ts) < 200):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    t_layer = ts[len(ts) // 2]
    return t_layer, len(ts)


def run_reference(a):
    """--impl reference: rank 0 only; each STEP is a bounded sample (one layer of the 32) of the
    workload, timed on the host cores; reported tokens/s extrapolates to the full step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import hydragen_oracle as O

    cores = torch.get_num_threads()
    q, k, v, sk, sv, sl = cpu_layer_inputs(a, torch)
    run = lambda: O.hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl, compute_dtype=torch.float32)
    steps, warmup = min(a.steps, 20), min(a.warmup, 3)
    for _ in range(max(1, warmup)):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    t_layer = (time.perf_counter() - t0) / steps
    ms_step = t_layer * a.layers * 1e3
    value = a.batch / (ms_step / 1e3)
    sample = f"each step = 1 of {a.layers} layers of the workload (fp32 torch on {cores} host threads); tokens/s extrapolated x{a.layers}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "parallelism": "host cpu", "l2": "n/a (cpu)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of hydragen/attention.py:177-354 + flash.py semantics (oracle/hydragen_oracle.py); the reference has no CPU attention path and its CUDA path cannot import in this image (SURVEY.md 8c)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="hg_clocks_", suffix=".csv")
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            busy = sorted(x for x in sm if x > 0.5 * max(sm)) or sorted(sm)
            out.update(sm_mhz=busy[len(busy) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def _log(msg):
    if os.environ.get("HG_BENCH_VERBOSE"):
        print(f"[bench r{os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def run_ours(a):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit(f"--gpus {a.gpus} needs torchrun with {a.gpus} ranks (one process per GPU)")
        a.gpus = world
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=dev)

    from hydragen_b200 import _lib
    from hydragen_b200.attention import hydragen_attention_decode
    from hydragen_b200.flash import decode_attention_fused, prefix_attention_grouped

    _lib.load()  # raises if the CUDA extension is missing: no fallback
    assert a.heads % world == 0 and a.kv_heads % world == 0, "heads must divide over the ranks (hydragen/tp.py:43-46)"
    H, HKV, D, B, L = a.heads // world, a.kv_heads // world, a.head_dim, a.batch, a.layers
    hidden = a.heads * a.head_dim
    dt = torch.bfloat16
    torch.manual_seed(1234 + rank)
    mk = lambda *s: torch.randn(*s, device=dev, dtype=dt)
    # per-layer state, resident in HBM before the timed region: shared prefix KV, unique KV caches
    shared_k = [mk(1, a.prefix_len, HKV, D) for _ in range(L)]
    shared_v = [mk(1, a.prefix_len, HKV, D) for _ in range(L)]
    uniq = torch.randn(L, 2, B, a.max_unique_len, HKV, D, device=dev, dtype=dt)
    # per-layer step inputs: q and the new token's k, v (what q/k/v_proj + RoPE hand to the hot path)
    qs = [mk(B, 1, H, D) for _ in range(L)]
    kn = [mk(B, 1, HKV, D) for _ in range(L)]
    vn = [mk(B, 1, HKV, D) for _ in range(L)]
    pos = torch.full((B, 1), a.suffix_len - 1, device=dev, dtype=torch.int64)  # row of the new token
    seq_lens = pos[:, 0] + 1
    # the collective of the path: all-reduce(sum) of the row-parallel o_proj output [B, hidden], one per layer
    # (hydragen/tp.py:108-112) -- the library's NVLS kernel on symmetric memory where the platform has NVLink
    # multicast, else NCCL
    proj, nvls = None, None
    if world > 1:
        from hydragen_b200.collectives import make_all_reduce

        nvls = None if os.environ.get("HG_BENCH_NCCL") else make_all_reduce(L * (B * hidden * 2 + 256), dev)
        if nvls is not None:
            proj = [nvls.buffer((B, hidden), dt).zero_() for _ in range(L)]
        else:
            proj = [torch.zeros(B, hidden, device=dev, dtype=dt) for _ in range(L)]
    all_reduce = (lambda t: nvls.all_reduce_(t)) if nvls is not None else (lambda t: dist.all_reduce(t))
    outs = [None] * L

    def layer(i):
        # prefix launch (tcgen05) + ONE launch for KV append + suffix attention + combine
        outs[i] = hydragen_attention_decode(qs[i], kn[i], vn[i], pos, uniq[i, 0], uniq[i, 1], [shared_k[i]], [shared_v[i]])
        if world > 1:
            all_reduce(proj[i])  # the one collective per attention layer (hydragen/tp.py:108-112)

    def step_eager():
        for i in range(L):
            layer(i)

    launches_per_step = L * (2 + (1 if nvls is not None else 0))  # prefix + fused append/suffix/combine (+ our all-reduce kernel; NCCL's not counted)

    _log("inputs ready; eager warm-up")
    for _ in range(3):
        step_eager()
    torch.cuda.synchronize()
    _log("eager warm-up done; capturing")
    graph = None
    if not a.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step_eager()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step_eager()
        except Exception as ex:  # NCCL capture can be refused on some setups: time eager launches instead
            graph = None
            torch.cuda.synchronize()
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({ex!r}); timing eager launches", file=sys.stderr)
    step = graph.replay if graph is not None else step_eager
    _log(f"capture done (graph={graph is not None})")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, a.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    t_ms = e0.elapsed_time(e1)
    _log(f"timed region done: {t_ms / a.steps:.3f} ms/step")
    clocks = sampler.stop() if sampler is not None else None
    if world > 1:
        tt = torch.tensor([t_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt.item())
    ms_step = t_ms / a.steps
    value = B / (ms_step / 1e3)

    # ---- roofline of the dominant kernel, measured live: events around every prefix launch inside full steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = float(peaks.get("bf16_tflops", 1590.0))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json, burst)" if peaks else "fallback (B200_PROFILING.md)"
    # Per-kernel time, live: a CUDA graph holding ONLY that kernel's launches of one step (one per layer, each
    # on its own layer's tensors so nothing is L2-warm from a previous launch), replayed and timed with CUDA
    # events on the launching stream.  (Events around eager launches would time the Python launch gaps.)
    pre_out = [None] * L

    from hydragen_b200.flash import prefix_attention_partials

    def only_prefix():  # the same launch the step makes: split-KV when the local heads alone do not fill the SMs (TP ranks)
        for i in range(L):
            pre_out[i] = prefix_attention_partials(qs[i], shared_k[i], shared_v[i], n_groups=1, max_splits=_lib.HG_MAX_COMBINE)

    def only_suffix():
        for i in range(L):
            decode_attention_fused(qs[i], kn[i], vn[i], pos, uniq[i, 0], uniq[i, 1], pre_out[i][0], pre_out[i][1])

    def time_kernel_graph(fn, reps=20):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            g.replay()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) * 1e3 / (reps * L)  # us per launch

    # the step right before the path: RoPE of the new q / k rows, one launch per layer, in place
    from hydragen_b200.rope import apply_rotary_pos_emb

    max_pos = 4096
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2, device=dev, dtype=torch.float32) / D))
    ang = torch.outer(torch.arange(max_pos, device=dev, dtype=torch.float32), inv)
    cos_t, sin_t = torch.cat((ang, ang), -1).cos().to(dt), torch.cat((ang, ang), -1).sin().to(dt)
    abs_pos = torch.full((B, 1), a.prefix_len + a.suffix_len - 1, device=dev, dtype=torch.int64)
    rq = [t.clone() for t in qs]
    rk = [t.clone() for t in kn]

    def only_rope():
        for i in range(L):
            apply_rotary_pos_emb(rq[i], rk[i], cos_t, sin_t, abs_pos, inplace=True)

    pre_t = time_kernel_graph(only_prefix)
    suf_t = time_kernel_graph(only_suffix)
    rope_t = time_kernel_graph(only_rope)
    del rq, rk
    _log(f"per-kernel timing done: prefix {pre_t:.1f} us, suffix {suf_t:.1f} us, rope {rope_t:.1f} us")
    pre_flops = 4.0 * B * H * a.prefix_len * D
    esz = 2
    # new K,V rows read + appended, older K,V rows read, q + prefix partial + out, 2 LSE rows
    n_part = len(pre_out[0][0])  # prefix partials merged by the decode launch (1 unless split-KV)
    suf_bytes = (4.0 * B * HKV * D * esz + 2.0 * B * (a.suffix_len - 1) * HKV * D * esz + (2.0 + n_part) * B * H * D * esz
                 + (1.0 + n_part) * B * H * 4)
    roofline = {"kernel": "prefix_attn_sm100_kernel (tcgen05)", "bound": "tensor", "achieved": pre_flops / pre_t / 1e6, "peak": tf_peak,
                "unit": "TFLOP/s", "frac": pre_flops / pre_t / 1e6 / tf_peak, "traffic": None, "us_per_launch": pre_t,
                "algorithmic_flop_per_launch": pre_flops, "peak_source": peak_src}
    roofline_suffix = {"kernel": "decode_slot_kernel (kv append + suffix + combine)", "bound": "hbm", "achieved": suf_bytes / suf_t / 1e3, "peak": hbm_peak,
                       "unit": "GB/s", "frac": suf_bytes / suf_t / 1e3 / hbm_peak, "traffic": None, "us_per_launch": suf_t,
                       "algorithmic_bytes_per_launch": suf_bytes, "peak_source": peak_src}
    rope_bytes = 2.0 * B * (H + HKV) * D * esz  # q and k rows read and written once (table rows stay in cache)
    roofline_rope = {"kernel": "rope_qk_kernel (next row N2: RoPE of the new q/k rows, not part of the timed step)", "bound": "hbm",
                     "achieved": rope_bytes / rope_t / 1e3, "peak": hbm_peak, "unit": "GB/s", "frac": rope_bytes / rope_t / 1e3 / hbm_peak,
                     "traffic": None, "us_per_launch": rope_t, "algorithmic_bytes_per_launch": rope_bytes, "peak_source": peak_src}
    prof = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch read from the committed ncu --set full capture
    if os.path.exists(prof):
        try:
            tr = json.load(open(prof))
            roofline["traffic"] = tr.get("prefix_dram_bytes_per_launch")
            roofline_suffix["traffic"] = tr.get("suffix_dram_bytes_per_launch")
            roofline_rope["traffic"] = tr.get("rope_dram_bytes_per_launch")
        except Exception:
            pass

    # ---- e2e: host buffers, H2D of every layer's step inputs and D2H of its result inside the timed region
    e2e = None
    if a.e2e_steps > 0:
        hq = [torch.randn(B, 1, H, D, dtype=dt).pin_memory() for _ in range(L)]
        hk = [torch.randn(B, 1, HKV, D, dtype=dt).pin_memory() for _ in range(L)]
        hv = [torch.randn(B, 1, HKV, D, dtype=dt).pin_memory() for _ in range(L)]
        ho = [torch.empty(B, 1, H, D, dtype=dt).pin_memory() for _ in range(L)]

        from hydragen_b200.host import HostDecodeLayer, HostDecodePipeline

        pipe = HostDecodePipeline(dev)
        host_layers = [HostDecodeLayer(hq[i], hk[i], hv[i], ho[i], qs[i], kn[i], vn[i], uniq[i, 0], uniq[i, 1], [shared_k[i]], [shared_v[i]])
                       for i in range(L)]
        after = (lambda i: all_reduce(proj[i])) if world > 1 else None

        def step_e2e():
            pipe.step(host_layers, pos, after_layer=after)

        for _ in range(3):
            step_e2e()
        pipe.synchronize()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(a.e2e_steps):
            step_e2e()
        compute_stream = torch.cuda.current_stream()
        compute_stream.wait_stream(pipe.h2d)
        compute_stream.wait_stream(pipe.d2h)  # the last download is inside the timed region
        f1.record()
        barrier()
        te = f0.elapsed_time(f1)
        if world > 1:
            tt = torch.tensor([te], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt.item())
        h2d = L * (B * H * D + 2 * B * HKV * D) * esz * world
        d2h = L * B * H * D * esz * world
        e2e = {"value": B / (te / a.e2e_steps / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": te / a.e2e_steps, "steps": a.e2e_steps,
               "api": "hydragen_b200.host.HostDecodePipeline.step: pinned host q/k_new/v_new -> H2D -> kernels -> D2H of out, per layer, 3 streams"}

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        t_layer, n = time_cpu_layer(a, torch, a.cpu_seconds)
        cpu_baseline = {"value": B / (t_layer * L), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"1 of {L} layers of the same workload, fp32 torch, median of {n} runs (~{a.cpu_seconds:.0f} s); tokens/s extrapolated x{L}",
                        "ms_per_layer": t_layer * 1e3}

    used_graph = graph is not None
    full_model = None
    if a.full_model and world == 1:
        del graph, step
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import full_model_decode

        full_model = full_model_decode.run()
        graph = step = None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(a), "scope": "attention hot path of one decode step (prefix + fused kv-append/suffix/combine per layer); projections/MLP/sampling out of scope",
                       "parallelism": (f"tp{world} (head axis, 1 all-reduce of [B,{hidden}] bf16 per layer: " + ("hg_allreduce_multimem NVLS kernel" if nvls is not None else "NCCL") + ")") if world > 1 else "single GPU",
                       "l2": f"inputs larger than L2: {L} layers x distinct caches cycle {L * (2 * a.prefix_len * HKV * D * 2 + 4 * B * H * D * 2) / 2**20:.0f}+ MiB per step through a 126 MB L2",
                       "cuda_graph": used_graph},
            "roofline": roofline, "roofline_suffix": roofline_suffix, "roofline_rope": roofline_rope, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches_per_step * a.steps, "gpu_launches_per_step": launches_per_step, "clocks": clocks,
        }
        if full_model is not None:
            line["full_model"] = full_model
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL teardown can block while captured graphs still hold its kernels: drop them first, and never let
        # a stuck teardown outlive the printed result
        import threading

        threading.Timer(20.0, lambda: os._exit(0)).start()
        del step, graph
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
