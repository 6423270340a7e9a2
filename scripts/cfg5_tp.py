"""BASELINE.json configs[4] (torchrun, one rank per GPU): random-init Llama-2-13B sharded in memory along the head axis
(hydragen/tp.py:30-124 -> 5 heads per GPU at tp = 8), ONE shared prefix of 16384 tokens, 2048 completions x 256 new tokens,
bf16, CUDA-graph decode -- the setting of the reference's docs/sweeps_from_paper.md:35-41 -- measured the way
scripts/synth.py does (generate(N) minus generate(1), CUDA events, max over ranks).  Also reported, from the same
process: the attention share (disable_attention ablation), and the three kernels of the hot path timed alone on this
rank's shapes (prefix launch, fused append/suffix/combine launch at mid-decode, the [B, hidden] all-reduce).
CFG5_PREFIX / CFG5_BATCH / CFG5_NEW / CFG5_LAYERS shrink it for a smoke run.  Rank 0 prints one JSON object."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200 import _lib  # noqa: E402
from hydragen_b200.flash import decode_attention_fused, prefix_attention_partials  # noqa: E402
from hydragen_b200.llama import llama_config  # noqa: E402
from hydragen_b200.tp import _AllReduce, from_config_tp  # noqa: E402

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
PREFIX, BATCH, NEW = int(os.environ.get("CFG5_PREFIX", "16384")), int(os.environ.get("CFG5_BATCH", "2048")), int(os.environ.get("CFG5_NEW", "256"))
over = {"max_position_embeddings": max(4096, PREFIX + NEW + 16)}  # synthetic weights: the RoPE table simply covers the run
if os.environ.get("CFG5_LAYERS"):
    over["num_hidden_layers"] = int(os.environ["CFG5_LAYERS"])
cfg = llama_config("llama-2-13b", **over)
heads_total = cfg.num_attention_heads
t0 = time.time()
model = from_config_tp(cfg, dtype=torch.bfloat16, device=dev, seed=0)
model.setup_caches(max_unique_batch_size=BATCH, max_unique_seq_length=NEW, max_shared_batch_sizes=[1], max_shared_seq_lengths=[PREFIX])
model.graph(True)
ids = torch.randint(3, 31000, (1, PREFIX), device=dev)


def timed_generate(n_new, **kw):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.generate(input_ids=ids, num_return_sequences=BATCH, max_new_tokens=n_new, temperature=100.0, **kw)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


timed_generate(4)  # warm-up: graph capture, cuBLAS handles, NVLS arena
setup_s = time.time() - t0
full = timed_generate(NEW)
pre = timed_generate(1)
noattn = timed_generate(NEW, disable_attention=True)
noattn_pre = timed_generate(1, disable_attention=True)
steps = NEW - 1
dec_ms, na_ms = full - pre, noattn - noattn_pre
res = {"config": f"Llama-2-13B random init, tp={world} ({heads_total // world} heads/GPU), shared prefix {PREFIX}, batch {BATCH}, {NEW} new tokens, bf16, CUDA-graph decode",
       "layers": cfg.num_hidden_layers, "setup_s": round(setup_s, 1), "decode_tokens_per_s": BATCH * steps / (dec_ms / 1e3), "ms_per_decode_step": dec_ms / steps,
       "prefill_ms": pre, "no_attention_tokens_per_s": BATCH * steps / (na_ms / 1e3), "attention_share": max(0.0, 1 - na_ms / dec_ms),
       "nvls_all_reduce": any(v is not None for v in _AllReduce._arenas.values())}

# ---- the kernels of the path alone, on this rank's shapes -------------------------------------------------------
H = heads_total // world
D = 128
NL = 4
mk = lambda *s: torch.randn(*s, device=dev, dtype=torch.bfloat16)
q = [mk(BATCH, 1, H, D) for _ in range(NL)]
sk, sv = [mk(1, PREFIX, H, D) for _ in range(NL)], [mk(1, PREFIX, H, D) for _ in range(NL)]
kn, vn = [mk(BATCH, 1, H, D) for _ in range(NL)], [mk(BATCH, 1, H, D) for _ in range(NL)]
kc, vc = [mk(BATCH, NEW, H, D) for _ in range(NL)], [mk(BATCH, NEW, H, D) for _ in range(NL)]
pos = torch.full((BATCH,), NEW // 2 - 1, device=dev, dtype=torch.int64)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timed(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / NL)
    ts.sort()
    t = torch.tensor([ts[len(ts) // 2]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


parts = [None] * NL


def f_prefix():
    for i in range(NL):
        parts[i] = prefix_attention_partials(q[i], sk[i], sv[i], 1, max_splits=_lib.HG_MAX_COMBINE)


t_pre = timed(f_prefix)
t_suf = timed(lambda: [decode_attention_fused(q[i], kn[i], vn[i], pos, kc[i], vc[i], parts[i][0], parts[i][1]) for i in range(NL)])
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
flops = 4.0 * BATCH * H * PREFIX * D
n_part = len(parts[0][0])
suf_bytes = 4.0 * BATCH * H * D * 2 + 2.0 * BATCH * (NEW // 2 - 1) * H * D * 2 + 2.0 * BATCH * H * D * 2 + BATCH * H * 4
res["kernels_per_layer"] = {
    "prefix_us": t_pre, "prefix_kv_splits": n_part, "prefix_tflops": flops / t_pre / 1e6, "prefix_frac_of_tensor_peak": flops / t_pre / 1e6 / float(peaks.get("bf16_tflops", 1590.0)),
    "fused_suffix_us_at_mid_decode": t_suf, "fused_suffix_gbs": suf_bytes / t_suf / 1e3, "fused_suffix_frac_of_hbm_peak": suf_bytes / t_suf / 1e3 / float(peaks.get("hbm_gbs", 6650.0)),
}
if world > 1:
    from hydragen_b200.collectives import make_all_reduce

    msg = BATCH * cfg.hidden_size * 2
    ar = make_all_reduce(NL * (msg + 256), dev)
    if ar is not None:
        bufs = [ar.buffer((BATCH, cfg.hidden_size), torch.bfloat16).zero_() for _ in range(NL)]
        t_ar = timed(lambda: [ar.all_reduce_(b) for b in bufs])
    plain = [torch.zeros(BATCH, cfg.hidden_size, device=dev, dtype=torch.bfloat16) for _ in range(NL)]
    t_nccl = timed(lambda: [dist.all_reduce(p) for p in plain])
    res["all_reduce"] = {"message_mib": msg / 2**20, "nvls_kernel_us": t_ar if ar is not None else None, "nccl_us": t_nccl,
                         "per_decode_step_ms_nvls": (2 * cfg.num_hidden_layers * t_ar / 1e3) if ar is not None else None,
                         "share_of_decode_step": (2 * cfg.num_hidden_layers * t_ar / 1e3) / (dec_ms / steps) if ar is not None else None,
                         "note": "two all-reduces per decoder layer (attention o_proj + MLP down_proj, hydragen/tp.py:83-87, 108-112)"}
res["attention_per_decode_step_ms_estimate"] = cfg.num_hidden_layers * (t_pre + t_suf) / 1e3
if rank == 0:
    print(json.dumps(res), flush=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
os._exit(0)
