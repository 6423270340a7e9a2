"""Multi-GPU check + timing (torchrun, >= 2 GPUs) of the o_proj GEMM fused with its all-reduce (csrc/oproj_allreduce.cu,
SURVEY §8 N4; hydragen/llama.py:592-594 + hydragen/tp.py:108-112):
  parity   fused launch == torch matmul per rank (fp32 accumulate, 16-bit output) summed with NCCL, on several shapes, repeated
           calls on the same buffer (epoch flags), and replayed from a CUDA graph;
  timing   per (GEMM, all-reduce) pair, graph-timed, max over ranks: cuBLAS GEMM + NCCL, cuBLAS GEMM into the symmetric
           buffer + the stand-alone NVLS kernel (what tp.py runs today), and the fused launch.
OPROJ_SHAPES="m,n,hidden_k;..." (k is the FULL reduction length, split over the ranks); OPROJ_CTAS = CTAs for the fused launch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200.collectives import MultimemAllReduce  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
NL = 8
dt = torch.bfloat16
shapes = [tuple(int(v) for v in s.split(",")) for s in os.environ.get("OPROJ_SHAPES", "1024,4096,4096;2048,5120,5120;200,1032,1024").split(";")]
n_ctas = int(os.environ.get("OPROJ_CTAS", "0"))
big = max(m * n for m, n, _ in shapes)
ar = MultimemAllReduce(NL * (big * 2 + 256) + 4096, dev)
assert ar.available
ok = True


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / (reps * NL)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for m, n, kfull in shapes:
    k = kfull // world
    ar._used = 0
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    xs = [torch.randn(m, k, device=dev, generator=g).to(dt) for _ in range(NL)]
    ws = [(torch.randn(n, k, device=dev, generator=g) / kfull**0.5).to(dt) for _ in range(NL)]
    bufs = [ar.buffer((m, n), dt) for _ in range(NL)]
    plain = [torch.empty(m, n, device=dev, dtype=dt) for _ in range(NL)]
    # reference: per-rank product rounded to the 16-bit type (as any GEMM would store it), summed in fp32 over the ranks
    refs = []
    for x, w in zip(xs, ws):
        r = (x.float() @ w.float().t()).to(dt).float()
        dist.all_reduce(r)
        refs.append(r)
    scale = max(r.abs().max().item() for r in refs)

    def check(tag):
        global ok
        torch.cuda.synchronize()
        err = max((b.float() - r).abs().max().item() for b, r in zip(bufs, refs))
        good = err <= scale / 64  # bf16 rounding of the sum (2^-8 relative) + of each partial
        ok = ok and good
        if rank == 0:
            print(f"  [{m},{n}] k={k}/rank {tag}: max |diff| {err:.3e} (max |ref| {scale:.2f}) {'ok' if good else 'FAILED'}", flush=True)

    def fused():
        for x, w, b in zip(xs, ws, bufs):
            ar.linear_all_reduce_(x, w, b, n_ctas)

    for b in bufs:
        b.fill_(float("nan"))
    dist.barrier()
    fused()
    check("first call")
    for _ in range(3):
        fused()  # the same buffers again: partials overwrite the previous sums
    check("repeated")
    # stress: many back-to-back eager rounds on the same buffers, checked after each (a stale read of a partial shows as the
    # previous sum leaking into the new one)
    bad = 0
    for it in range(int(os.environ.get("OPROJ_STRESS", "0"))):
        fused()
        torch.cuda.synchronize()
        err = max((b.float() - r).abs().max().item() for b, r in zip(bufs, refs))
        bad += int(not err <= scale / 64)
    if os.environ.get("OPROJ_STRESS"):
        t = torch.tensor([bad], device=dev)
        dist.all_reduce(t)
        ok = ok and int(t.item()) == 0
        if rank == 0:
            print(f"  [{m},{n}] stress: {int(t.item())} bad rounds (summed over ranks) of {os.environ['OPROJ_STRESS']}", flush=True)
    t_fused = timed(fused)
    check("graph replay")
    if os.environ.get("OPROJ_FUSED_ONLY"):
        if rank == 0:
            print(f"world {world} [{m},{n}] k={k}/rank, us per o_proj + all-reduce: fused {t_fused:.1f}", flush=True)
        continue

    def lib_nvls():
        for x, w, b in zip(xs, ws, bufs):
            torch.matmul(x, w.t(), out=b)
            ar.all_reduce_(b)

    def lib_nccl():
        for x, w, p in zip(xs, ws, plain):
            torch.matmul(x, w.t(), out=p)
            dist.all_reduce(p)

    def gemm_only():
        for x, w, p in zip(xs, ws, plain):
            torch.matmul(x, w.t(), out=p)

    def ar_only():
        for b in bufs:
            ar.all_reduce_(b)

    t_nvls, t_nccl, t_gemm, t_ar = timed(lib_nvls), timed(lib_nccl), timed(gemm_only), timed(ar_only)
    if rank == 0:
        print(f"world {world} [{m},{n}] k={k}/rank, us per o_proj + all-reduce: fused {t_fused:.1f} | cuBLAS + NVLS kernel {t_nvls:.1f} "
              f"(GEMM alone {t_gemm:.1f}, all-reduce alone {t_ar:.1f}) | cuBLAS + NCCL {t_nccl:.1f}", flush=True)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("oproj_allreduce parity", "ok" if ok else "FAILED", flush=True)
os._exit(0 if ok else 1)
