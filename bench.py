#!/usr/bin/env python
"""bench.py -- decode throughput of the Hydragen shared-prefix attention hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the microbenchmark shape, for every layer of Llama-2-7B):
B = 1024 sequences decoding one token each against ONE shared prefix of 2048 tokens, every
sequence owning `--suffix-len` tokens of its own KV (default 1), 32 query / 32 kv heads, d = 128,
bf16.  One STEP = the ATTENTION hot path of one whole-model decode step: for each of the 32 layers
(each with its OWN caches, so 1.5 GB+ of distinct inputs stream through the 126 MB L2 per step)
    prefix launch (tcgen05)  ->  ONE launch: KV append of the new token + suffix attention + combine
i.e. what the reference's decode branch does with update_per_completion_kvs + hydragen_attention
(hydragen/llama.py:564-587), through ``hydragen_attention_decode``.
The step is captured in a CUDA graph (the reference replays graphs too: llama.py:781-866,
benchmark_utils.py:140-170).  `value` = B / step time = decode tokens/s of the ATTENTION path at the
lightest point of a decode (suffix 1); the projections / MLP / sampling around it are out of scope
(SURVEY.md section 8).  What the headline leaves out is reported beside it, measured in the same run:
  `suffix_sweep`       the fused append/suffix/combine launch at suffix 1 / 32 / 64 / 127 (the dominant kernel of a real
                       decode), HBM fraction on the bytes that must cross HBM;
  `decode_integrated`  attention time of an average step of BASELINE.json configs[2] (suffix 1..127), from the sweep;
  `full_model`         configs[2] itself: random-init Llama-2-7B generate(1024 x 128 tokens), whole-model tokens/s;
  `sustained`          the same step replayed for seconds, against the sustained cuBLAS peak, with clocks;
  `lib_fa2`            the installed flash-attn 2.x kernels on the same tensors (the library the reference calls);
  `oproj_allreduce`    N > 1: the o_proj GEMM + all-reduce pair per layer, library GEMM + NVLS kernel vs ONE fused launch, alone and in the step;
  `hierarchy_cfg4`     configs[3]: the two-level hierarchy 1 x 1024 -> 32 x 64 -> B = 1024, one grouped prefix launch per layer.

Multi-GPU = the reference's head-axis tensor parallelism (hydragen/tp.py): each rank runs the same
step on Hq/N local heads, then per layer ONE all-reduce of the [B, hidden] bf16 tensor that
the row-parallel o_proj would produce.  Total work is fixed -> "scaling": "strong".

`e2e`: the same step through hydragen_b200.host.HostDecodePipeline: every layer's q / k_new / v_new come
from pinned HOST memory and the attention output is read back to the host inside the timed region
(one CUDA graph per step: copies and kernels on three streams).
`--impl reference`: the CPU restatement of the reference's algorithm (oracle/, torch fp32 on all
host cores; the reference itself has no CPU attention path and cannot run here: DESIGN.md).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "decode tokens/sec (Llama-2-7B, 2K shared prefix, bs=1024), attention hot path"
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
    ap.add_argument("--sustain-seconds", type=float, default=2.0, help="length of the sustained arm (0: skip)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the suffix-length sweep / decode-integrated figures")
    ap.add_argument("--no-lib", action="store_true", help="skip the flash-attn library timings")
    ap.add_argument("--no-full-model", action="store_true", help="skip the whole-model generate of configs[2]")
    ap.add_argument("--full-model", action="store_true", help="(default at N = 1; kept for compatibility)")
    return ap.parse_args()


def workload_name(a):
    return (f"microbenchmark cfg#2 x {a.layers} layers: B={a.batch}, shared prefix {a.prefix_len}, suffix {a.suffix_len}, "
            f"{a.heads}q/{a.kv_heads}kv heads d={a.head_dim}, bf16")


def bench_config(a, world):
    """The same dict in both arms (computed from the arguments only)."""
    hidden = a.heads * a.head_dim
    per_layer = 2 * a.prefix_len * a.kv_heads * a.head_dim * 2 + 4 * a.batch * a.heads * a.head_dim * 2
    return {
        "workload": workload_name(a),
        "scope": "attention hot path of one decode step (tcgen05 prefix launch + fused kv-append/suffix/combine launch per layer); "
                 "projections/MLP/sampling out of scope; whole-model numbers under full_model",
        "parallelism": f"tp{world} (head axis, 1 all-reduce of [B,{hidden}] bf16 per layer)" if world > 1 else "single GPU",
        "l2": f"inputs larger than L2: {a.layers} layers x distinct caches cycle {a.layers * per_layer / 2**20:.0f}+ MiB per step through a 126 MB L2",
        "cuda_graph": not a.no_graph,
    }


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle = restatement of the reference's algorithm; test/bench infrastructure only)
# ------------------------------------------------------------------------------------------------


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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

    torch.set_num_threads(host_threads())
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
    """--impl reference: rank 0 only, all host threads (torchrun exports OMP_NUM_THREADS=1: overridden here).  Runs
    EXACTLY --warmup + --steps steps.  A step is the whole workload -- `layers` layer-calls of the restated
    hydragen_attention -- when K such steps fit ~150 s; otherwise a bounded sample of it (fewer layer-calls per step,
    said in `sample`, tokens/s scaled to the full step).  ms_per_step is what a step really took."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import hydragen_oracle as O

    cores = host_threads()
    torch.set_num_threads(cores)
    q, k, v, sk, sv, sl = cpu_layer_inputs(a, torch)
    run = lambda: O.hydragen_attention_nopad(q, k, v, [sk], [sv], seq_len=sl, compute_dtype=torch.float32)
    run()
    t0 = time.perf_counter()
    run()
    t_probe = time.perf_counter() - t0
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    budget = 150.0
    per_step = max(1, min(a.layers, int(budget / ((steps + warmup) * t_probe))))
    for _ in range(warmup):
        for _ in range(per_step):
            run()
    t0 = time.perf_counter()
    for _ in range(steps):
        for _ in range(per_step):
            run()
    t_step = (time.perf_counter() - t0) / steps
    ms_full = t_step / per_step * a.layers * 1e3
    value = a.batch / (ms_full / 1e3)
    sample = (f"each step = {per_step} of the {a.layers} layer-calls of the workload on one layer's tensors (fp32 torch, {cores} host threads)"
              + ("" if per_step == a.layers else f"; tokens/s scaled x{a.layers / per_step:.1f} to the full step"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": t_step * 1e3, "ms_per_full_step": ms_full, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": bench_config(a, a.gpus),
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
    """SM clock, power and throttle reasons of one GPU sampled through NVML from a thread (~1 ms period: the timed
    region of the default run is tens of milliseconds, far below nvidia-smi's sampling period)."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, torch, cuda_index):
        self.ok = False
        self.sm, self.power, self.mask = [], [], 0
        self.sm_max = None
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[cuda_index]) if vis and vis.split(",")[cuda_index].isdigit() else cuda_index
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv, self.h = pynvml, h
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False
        self.thread = None

    def _loop(self):
        nv, h = self.nv, self.h
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(get_reasons(h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3)
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.ok:
            self.sm, self.power, self.mask = [], [], 0
            self._stop.clear()
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        return self

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "power_w_max": None, "source": "nvml thread, ~1 ms period"}
        if not self.ok or self.thread is None:
            return out
        self._stop.set()
        self.thread.join(timeout=2)
        if self.sm:
            s = sorted(self.sm)
            out.update(sm_mhz=s[len(s) // 2], sm_mhz_min=s[0], samples=len(s), power_w_max=max(self.power) if self.power else None,
                       reasons=sorted(n for bit, n in self.REASONS.items() if self.mask & bit))
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
    numa = None
    if world > 1:
        # one process per GPU: keep it -- and the pinned host buffers of the e2e leg, allocated below -- on the socket its GPU hangs
        # off (r02x / r02zk: without placement the aggregate host->device rate stops at ~115 GB/s from N = 4 on).  Not at N = 1: the
        # cpu_baseline leg of that run uses every host core.
        from hydragen_b200.host import bind_process_to_gpu_numa

        numa = bind_process_to_gpu_numa(local_rank)
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
    esz = 2
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

    def make_graph(fn):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            tt = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return x

    _log("inputs ready; eager warm-up")
    for _ in range(3):
        step_eager()
    torch.cuda.synchronize()
    _log("eager warm-up done; capturing")
    graph = None
    if not a.no_graph:
        try:
            graph = make_graph(step_eager)
        except Exception as ex:  # NCCL capture can be refused on some setups: time eager launches instead
            graph = None
            torch.cuda.synchronize()
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({ex!r}); timing eager launches", file=sys.stderr)
    step = graph.replay if graph is not None else step_eager
    _log(f"capture done (graph={graph is not None})")

    warm = max(3, a.warmup)
    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(torch, local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler is not None:
        sampler.start()
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler is not None else None
    t_ms = max_over_ranks(e0.elapsed_time(e1))
    _log(f"timed region done: {t_ms / a.steps:.3f} ms/step")
    ms_step = t_ms / a.steps
    value = B / (ms_step / 1e3)

    # ---- peaks ---------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = float(peaks.get("bf16_tflops", 1590.0))
    tf_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json, burst)" if peaks else "fallback (B200_PROFILING.md)"

    # ---- sustained arm: the same step replayed for seconds (power-capped clocks), clocks sampled -----------------
    sustained = None
    if a.sustain_seconds > 0 and graph is not None:
        n = max(a.steps, int(a.sustain_seconds * 1e3 / ms_step))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sm2 = ClockSampler(torch, local_rank).start() if rank == 0 else None
        s0.record()
        for _ in range(n):
            step()
        s1.record()
        barrier()
        ck = sm2.stop() if sm2 is not None else None
        ts = max_over_ranks(s0.elapsed_time(s1)) / n
        sustained = {"value": B / (ts / 1e3), "unit": UNIT, "ms_per_step": ts, "steps": n, "seconds": ts * n / 1e3, "clocks": ck}

    # ---- per-kernel time, live: a CUDA graph holding ONLY that kernel's launches of one step (one per layer, each on
    # its own layer's tensors so nothing is L2-warm from a previous launch), replayed and timed with CUDA events on the
    # launching stream.  (Events around eager launches would time the Python launch gaps.)
    pre_out = [None] * L

    from hydragen_b200.flash import prefix_attention_partials

    def only_prefix():  # the same launch the step makes: split-KV when the local heads alone do not fill the SMs (TP ranks)
        for i in range(L):
            pre_out[i] = prefix_attention_partials(qs[i], shared_k[i], shared_v[i], 1, max_splits=_lib.HG_MAX_COMBINE)

    def only_suffix():
        for i in range(L):
            decode_attention_fused(qs[i], kn[i], vn[i], pos, uniq[i, 0], uniq[i, 1], pre_out[i][0], pre_out[i][1])

    def time_graph(g, launches, reps=20, seconds=0.0):
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            g.replay()
        t1.record()
        torch.cuda.synchronize()
        us = t0.elapsed_time(t1) * 1e3 / (reps * launches)
        if seconds > 0:  # again, for `seconds`: the power-capped figure
            reps2 = max(reps, int(seconds * 1e6 / (us * launches)))
            t0.record()
            for _ in range(reps2):
                g.replay()
            t1.record()
            torch.cuda.synchronize()
            return us, t0.elapsed_time(t1) * 1e3 / (reps2 * launches)
        return us, None

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

    g_pre = make_graph(only_prefix)
    pre_t, pre_t_sust = time_graph(g_pre, L, seconds=min(1.0, a.sustain_seconds))
    suf_t, _ = time_graph(make_graph(only_suffix), L)
    rope_t, _ = time_graph(make_graph(only_rope), L)
    del rq, rk, g_pre
    _log(f"per-kernel timing done: prefix {pre_t:.1f} us, suffix {suf_t:.1f} us, rope {rope_t:.1f} us")
    pre_flops = 4.0 * B * H * a.prefix_len * D

    def suffix_bytes(t):
        """(bytes that must cross HBM, bytes incl. the prefix partial the launch re-reads -- from L2 when it runs right
        behind the prefix launch) of the fused append/suffix/combine launch at suffix length t"""
        hbm = 4.0 * B * HKV * D * esz + 2.0 * B * (t - 1) * HKV * D * esz + 2.0 * B * H * D * esz + B * H * 4
        return hbm, hbm + n_part * (B * H * D * esz + B * H * 4)

    n_sms = _lib.load().hg_sm_count() or 148
    n_part = len(pre_out[0][0])  # prefix partials merged by the fused launch (1 unless split-KV)
    n_ctas = ((B + 255) // 256) * H * n_part
    roofline = {"kernel": "prefix_unit_sm100_kernel (tcgen05, one CTA per (256-row tile, head))", "bound": "tensor", "achieved": pre_flops / pre_t / 1e6, "peak": tf_peak,
                "unit": "TFLOP/s", "frac": pre_flops / pre_t / 1e6 / tf_peak, "traffic": None, "us_per_launch": pre_t,
                "algorithmic_flop_per_launch": pre_flops, "peak_source": peak_src, "grid_ctas": n_ctas, "kv_splits": n_part, "sms": n_sms}
    if pre_t_sust is not None:
        roofline["sustained"] = {"us_per_launch": pre_t_sust, "achieved": pre_flops / pre_t_sust / 1e6, "peak": tf_sust,
                                 "frac": pre_flops / pre_t_sust / 1e6 / tf_sust, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback"}
    sb_hbm, sb_l2 = suffix_bytes(a.suffix_len)
    roofline_suffix = {"kernel": "decode_slot_kernel (kv append + suffix + combine)", "bound": "hbm", "achieved": sb_hbm / suf_t / 1e3, "peak": hbm_peak,
                       "unit": "GB/s", "frac": sb_hbm / suf_t / 1e3 / hbm_peak, "traffic": None, "us_per_launch": suf_t,
                       "algorithmic_bytes_per_launch": sb_hbm, "frac_incl_l2_resident_prefix_partial": sb_l2 / suf_t / 1e3 / hbm_peak,
                       "note": f"suffix {a.suffix_len}: latency-bound at suffix 1 (the launch reads zero cache rows); see suffix_sweep for the sizes a decode spends its time at",
                       "peak_source": peak_src}
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

    # ---- suffix sweep: the fused append / suffix / combine launch at the suffix lengths a real decode runs at ------
    suffix_sweep, decode_integrated, lib_fa2 = None, None, None
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

    def time_flushed(g, launches, iters=12):
        ts = []
        for _ in range(iters):
            flush.zero_()  # 256 MiB > L2
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            g.replay()
            t1.record()
            torch.cuda.synchronize()
            ts.append(t0.elapsed_time(t1) * 1e3 / launches)
        ts.sort()
        return ts[len(ts) // 2]

    if not a.no_sweep:
        SL, MAXU = 4, 128  # 4 layers' worth of distinct [B, 128, Hkv, d] caches (2 GiB each at N = 1)
        big = torch.empty(SL, 2, B, MAXU, HKV, D, device=dev, dtype=dt).normal_()
        rows = []
        for t in (1, 32, 64, 127):
            p_t = torch.full((B, 1), t - 1, device=dev, dtype=torch.int64)

            def sweep_fn():
                for i in range(SL):
                    decode_attention_fused(qs[i], kn[i], vn[i], p_t, big[i, 0], big[i, 1], pre_out[i][0], pre_out[i][1])

            us = time_flushed(make_graph(sweep_fn), SL)
            hb, _ = suffix_bytes(t)
            rows.append({"suffix": t, "us_per_launch": us, "hbm_bytes": hb, "achieved_gbs": hb / us / 1e3, "frac_of_hbm_peak": hb / us / 1e3 / hbm_peak})
        suffix_sweep = {"kernel": "decode_slot_kernel", "rows": rows, "peak_gbs": hbm_peak,
                        "method": f"graph of {SL} launches on {SL} distinct caches of {MAXU} rows, L2 flushed (256 MiB) before every timed replay, median of 12"}
        # linear fit us(t) = c0 + c1 * t over the sweep -> mean over the 127 steps of configs[2] (t = 1 .. 127)
        n_ = len(rows)
        sx, sy = sum(r["suffix"] for r in rows), sum(r["us_per_launch"] for r in rows)
        sxx, sxy = sum(r["suffix"] ** 2 for r in rows), sum(r["suffix"] * r["us_per_launch"] for r in rows)
        c1 = (n_ * sxy - sx * sy) / (n_ * sxx - sx * sx)
        c0 = (sy - c1 * sx) / n_
        mean_suf = c0 + c1 * 64.0
        att_ms = L * (pre_t + mean_suf) / 1e3
        decode_integrated = {"attention_ms_per_step_mean": att_ms, "attention_only_tokens_per_s": B / (att_ms / 1e3),
                             "prefix_us": pre_t, "suffix_us_mean": mean_suf, "suffix_us_fit": {"c0": c0, "c1_per_token": c1},
                             "method": "mean over suffix t = 1..127 (BASELINE.json configs[2]: 128 new tokens) of layers x (prefix launch + fused launch), "
                                       "fused launch time from the linear fit of suffix_sweep; all-reduce not included"}
        del big
        _log("suffix sweep done")

    # ---- the library the reference calls, on the same tensors (flash-attn 2.x: mma.sync kernels recompiled for sm_100)
    if not a.no_lib and world == 1:
        try:
            import flash_attn
            from flash_attn import flash_attn_func, flash_attn_with_kvcache

            NLIB = 4
            sl32 = (pos[:, 0] + 1).to(torch.int32)

            def fa_prefix():
                for i in range(NLIB):
                    flash_attn_func(qs[i].view(1, B, H, D), shared_k[i], shared_v[i], softmax_scale=D**-0.5)

            def fa_suffix():
                for i in range(NLIB):
                    flash_attn_with_kvcache(qs[i], uniq[i, 0], uniq[i, 1], cache_seqlens=sl32, softmax_scale=D**-0.5)

            lib_fa2 = {"version": flash_attn.__version__, "prefix_us": time_flushed(make_graph(fa_prefix), NLIB),
                       "suffix_us": time_flushed(make_graph(fa_suffix), NLIB), "ours_prefix_us": time_flushed(make_graph(lambda: [prefix_attention_grouped(qs[i], shared_k[i], shared_v[i], n_groups=1) for i in range(NLIB)]), NLIB),
                       "method": f"graph of {NLIB} launches on distinct tensors, L2 flushed before every timed replay, median of 12; "
                                 "flash_attn_func on Q [1, B, H, d] x the shared prefix; flash_attn_with_kvcache on the unique cache (suffix length as benched); "
                                 "the combine the reference adds on top is not included"}
        except Exception as ex:
            lib_fa2 = {"unavailable": repr(ex)[:200]}
    del flush

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
        for _ in range(2):
            pipe.step(host_layers, pos, after_layer=after)
        pipe.synchronize()
        e2e_graph = None
        if not a.no_graph:
            try:
                e2e_graph = pipe.capture(host_layers, pos, after_layer=after)
            except Exception as ex:
                e2e_graph = None
                torch.cuda.synchronize()
                if rank == 0:
                    print(f"[bench] e2e graph capture failed ({ex!r}); issuing the pipeline eagerly", file=sys.stderr)
        step_e2e = e2e_graph.replay if e2e_graph is not None else (lambda: pipe.step(host_layers, pos, after_layer=after))
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
        te = max_over_ranks(f0.elapsed_time(f1))
        h2d = L * (B * H * D + 2 * B * HKV * D) * esz * world
        d2h = L * B * H * D * esz * world
        e2e = {"value": B / (te / a.e2e_steps / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": te / a.e2e_steps, "steps": a.e2e_steps, "cuda_graph": e2e_graph is not None,
               "pcie_gbs": {"h2d": h2d / world / (te / a.e2e_steps) / 1e6, "d2h": d2h / world / (te / a.e2e_steps) / 1e6},
               "api": "hydragen_b200.host.HostDecodePipeline (capture + replay): pinned host q/k_new/v_new -> H2D -> kernels -> D2H of out, per layer, 3 streams, one CUDA graph per step"}
        del e2e_graph, pipe, host_layers

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        t_layer, n = time_cpu_layer(a, torch, a.cpu_seconds)
        cpu_baseline = {"value": B / (t_layer * L), "unit": UNIT, "cores": host_threads(), "kind": "port",
                        "sample": f"1 of {L} layers of the same workload, fp32 torch, median of {n} runs (~{a.cpu_seconds:.0f} s); tokens/s extrapolated x{L}",
                        "ms_per_layer": t_layer * 1e3}

    # ---- BASELINE.json configs[3] at the operator level: two shared levels in one grouped prefix launch ----------
    hierarchy = None
    if world == 1 and not a.no_sweep:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        try:
            import time_hierarchy

            hierarchy = time_hierarchy.run(peaks=peaks or None)
        except Exception as ex:
            hierarchy = {"error": repr(ex)[:300]}

    # ---- the row-parallel o_proj GEMM + the all-reduce of its partials, per layer, as the library GEMM followed
    # by the stand-alone NVLS kernel (what `value` above contains is the collective alone) and as ONE fused launch
    # (csrc/oproj_allreduce.cu) -- the pair on its own and inside the step (attention -> o_proj -> all-reduce) ----------
    oproj = None
    if world > 1 and nvls is not None and not a.no_sweep:
        try:
            kloc = H * D
            ws = [torch.randn(hidden, kloc, device=dev, dtype=dt) / hidden**0.5 for _ in range(L)]
            xv = [torch.randn(B, kloc, device=dev, dtype=dt) for _ in range(L)]

            def attn(i):
                return hydragen_attention_decode(qs[i], kn[i], vn[i], pos, uniq[i, 0], uniq[i, 1], [shared_k[i]], [shared_v[i]]).view(B, kloc)

            def pair_fused():
                for i in range(L):
                    nvls.linear_all_reduce_(xv[i], ws[i], proj[i])

            def pair_lib():
                for i in range(L):
                    torch.matmul(xv[i], ws[i].t(), out=proj[i])
                    nvls.all_reduce_(proj[i])

            def step_fused():
                for i in range(L):
                    nvls.linear_all_reduce_(attn(i), ws[i], proj[i])

            def step_lib():
                for i in range(L):
                    torch.matmul(attn(i), ws[i].t(), out=proj[i])
                    nvls.all_reduce_(proj[i])

            def timed_multi(fn, launches):
                g = make_graph(fn)
                barrier()
                us, _ = time_graph(g, launches)
                del g
                return max_over_ranks(us)

            pf, pl = timed_multi(pair_fused, L), timed_multi(pair_lib, L)
            sf, sl = timed_multi(step_fused, 1) / 1e3, timed_multi(step_lib, 1) / 1e3
            oproj = {"what": f"per layer: o_proj [B={B}, {kloc}] x [{hidden}, {kloc}]^T (this rank's heads) + all-reduce of the [B, {hidden}] bf16 partials "
                             "(hydragen/llama.py:592-594 + tp.py:108-112); graph-timed, max over ranks",
                     "pair_us": {"fused_tcgen05_gemm_nvls": pf, "cublas_then_nvls_kernel": pl},
                     "step_ms": {"attention_then_fused": sf, "attention_then_cublas_then_nvls_kernel": sl},
                     "tokens_per_s": {"attention_then_fused": B / (sf / 1e3), "attention_then_cublas_then_nvls_kernel": B / (sl / 1e3)}}
            del ws, xv
        except Exception as ex:
            oproj = {"error": repr(ex)[:300]}

    used_graph = graph is not None
    full_model = None
    if world == 1 and not a.no_full_model:
        del graph, step
        qs = kn = vn = shared_k = shared_v = uniq = outs = pre_out = None
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        try:
            import full_model_decode

            full_model = full_model_decode.run(iters=1)
            full_model["what"] = ("BASELINE.json configs[2]: random-init Llama-2-7B, 1 shared prompt of 2048 tokens, 1024 completions x 128 new tokens, bf16, "
                                  "CUDA-graph decode; projections / MLP / lm_head are stock cuBLAS (out of scope of the hot path)")
        except Exception as ex:
            full_model = {"error": repr(ex)[:300]}
        graph = step = None

    if rank == 0:
        cfg = bench_config(a, world)
        cfg["cuda_graph"] = used_graph
        cfg["collective"] = ("hg_allreduce_multimem NVLS kernel" if nvls is not None else "NCCL") if world > 1 else None
        cfg["numa_bind_rank0"] = numa
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warm, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": cfg,
            "roofline": roofline, "roofline_suffix": roofline_suffix, "roofline_rope": roofline_rope, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": launches_per_step * a.steps, "gpu_launches_per_step": launches_per_step, "clocks": clocks,
            "sustained": sustained, "suffix_sweep": suffix_sweep, "decode_integrated": decode_integrated, "lib_fa2": lib_fa2,
            "hierarchy_cfg4": hierarchy, "oproj_allreduce": oproj,
        }
        if full_model is not None:
            line["full_model"] = full_model
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL teardown can block while captured graphs still hold its kernels: drop them first, and never let
        # a stuck teardown outlive the printed result
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
