"""Multi-GPU check of the NVLS all-reduce kernel (run under torchrun on >= 2 GPUs of one NVSwitch box):
result == NCCL all_reduce on the same inputs, then graph-timed latency of both at the decode-step message size."""
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
B, HID, NL = 1024, 4096, 8
ar = MultimemAllReduce(NL * B * HID * 2 + 4096, dev)
if rank == 0:
    print(f"multicast support: {ar.available} (world {world})", flush=True)
if not ar.available:
    dist.destroy_process_group()
    sys.exit(0)
bufs = [ar.buffer((B, HID), torch.bfloat16) for _ in range(NL)]
g = torch.Generator(device=dev).manual_seed(100 + rank)
ok = True
for it in range(3):
    src = [torch.randn(B, HID, device=dev, dtype=torch.bfloat16, generator=g) for _ in range(NL)]
    ref = [s.clone() for s in src]
    for b, s in zip(bufs, src):
        b.copy_(s)
    for b in bufs:
        ar.all_reduce_(b)
    for r in ref:
        dist.all_reduce(r)
    torch.cuda.synchronize()
    err = max((b.float() - r.float()).abs().max().item() for b, r in zip(bufs, ref))
    scale = max(r.float().abs().max().item() for r in ref)
    ok = ok and err <= 2e-2 * scale  # bf16 sums in a different order than NCCL's ring
    if rank == 0:
        print(f"iter {it}: max |nvls - nccl| = {err:.3e} (max |value| {scale:.2f})", flush=True)


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
    return e0.elapsed_time(e1) * 1e3 / (reps * NL)


t_nvls = timed(lambda: [ar.all_reduce_(b) for b in bufs])
# one-shot, out of place
for b, s_ in zip(bufs, src):
    b.copy_(s_)
outs1 = [torch.empty(B, HID, device=dev, dtype=torch.bfloat16) for _ in range(NL)]
for b, o in zip(bufs, outs1):
    ar.all_reduce(b, o)
torch.cuda.synchronize()
err1 = max((o.float() - r.float()).abs().max().item() for o, r in zip(outs1, ref))
ok = ok and err1 <= 2e-2 * scale
t_one = timed(lambda: [ar.all_reduce(b, o) for b, o in zip(bufs, outs1)])
if rank == 0:
    print(f"one-shot: max |nvls - nccl| = {err1:.3e}; {t_one:.1f} us", flush=True)
for nb in [int(x) for x in os.environ.get("AR_BLOCKS", "").split(",") if x]:
    ar2 = MultimemAllReduce(NL * B * HID * 2 + 4096, dev, n_blocks=nb)
    b2 = [ar2.buffer((B, HID), torch.bfloat16) for _ in range(NL)]
    t2 = timed(lambda: [ar2.all_reduce_(b) for b in b2])
    if rank == 0:
        print(f"  n_blocks {nb}: {t2:.1f} us", flush=True)
plain = [torch.zeros(B, HID, device=dev, dtype=torch.bfloat16) for _ in range(NL)]
t_nccl = timed(lambda: [dist.all_reduce(p) for p in plain])
if rank == 0:
    print(f"all-reduce of {B * HID * 2 / 2**20:.0f} MiB bf16 over {world} GPUs: NVLS kernel {t_nvls:.1f} us, NCCL {t_nccl:.1f} us; parity {'ok' if ok else 'FAILED'}", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0 if ok else 1)
