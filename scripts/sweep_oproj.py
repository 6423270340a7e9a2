"""Development aid (torchrun, >= 2 GPUs): the fused o_proj + all-reduce launch swept over its knobs in ONE process (they are read at
every launch): tile width, reductions per lane and slice (U), reduce warps per CTA.  Graph-timed, max over ranks; every
configuration is also checked against matmul + NCCL."""
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
NL, dt = 8, torch.bfloat16
shapes = [tuple(int(v) for v in s.split(",")) for s in os.environ.get("OPROJ_SHAPES", "1024,4096,4096;2048,5120,5120").split(";")]
big = max(m * n for m, n, _ in shapes)
ar = MultimemAllReduce(NL * (big * 2 + 256) + 4096, dev)
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
    ref = (xs[-1].float() @ ws[-1].float().t()).to(dt).float()
    dist.all_reduce(ref)
    scale = ref.abs().max().item()

    def fused():
        for x, w, b in zip(xs, ws, bufs):
            ar.linear_all_reduce_(x, w, b)

    def lib():
        for x, w, b in zip(xs, ws, bufs):
            torch.matmul(x, w.t(), out=b)
            ar.all_reduce_(b)

    t_lib = timed(lib)
    if rank == 0:
        print(f"world {world} [{m},{n}] k={k}/rank: cuBLAS + NVLS kernel {t_lib:.1f} us", flush=True)
    for bn in (128, 256):
        for u in (1, 2, 4):
            for rw in (1, 2, 4, 8):
                if bn == 256 and (u, rw) not in ((1, 4), (2, 2), (4, 2)):
                    continue
                os.environ.update(HYDRAGEN_B200_OPROJ_BN=str(bn), HYDRAGEN_B200_OPROJ_U=str(u), HYDRAGEN_B200_OPROJ_RWARPS=str(rw))
                t = timed(fused)
                err = (bufs[-1].float() - ref).abs().max().item()
                good = err <= scale / 64
                ok = ok and good
                if rank == 0:
                    print(f"  BN {bn} U {u} warps {rw} (in flight/GPU {148 * rw * 2 * u * 512 >> 10} KiB): fused {t:.1f} us{'' if good else '  PARITY FAILED'}", flush=True)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("sweep parity", "ok" if ok else "FAILED", flush=True)
os._exit(0 if ok else 1)
