"""Development aid (torchrun, >= 2 GPUs): graph-timed latency of hg_allreduce_multimem (two-shot in place, one-shot out
of place) and of NCCL at the decode-step message size, swept over the CTA count.  HYDRAGEN_B200_AR_UNROLL (4 | 8 | 16,
read once per process) = reductions in flight per thread; AR_BYTES_MIB = message size (default 8 = [1024, 4096] bf16)."""
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
mib = int(os.environ.get("AR_BYTES_MIB", "8"))
B, HID, NL = 1024, mib * 512, 8
ar = MultimemAllReduce(NL * B * HID * 2 + 4096, dev)
assert ar.available
bufs = [ar.buffer((B, HID), torch.bfloat16).zero_() for _ in range(NL)]
outs = [torch.empty(B, HID, device=dev, dtype=torch.bfloat16) for _ in range(NL)]
plain = [torch.zeros(B, HID, device=dev, dtype=torch.bfloat16) for _ in range(NL)]


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


unroll = os.environ.get("HYDRAGEN_B200_AR_UNROLL", "8")
for nb in [int(x) for x in os.environ.get("AR_BLOCKS", "16,32,64,128").split(",")]:
    ar.n_blocks = nb
    t = timed(lambda: [ar.all_reduce_(b) for b in bufs])
    t1 = timed(lambda: [ar.all_reduce(b, o) for b, o in zip(bufs, outs)]) if world <= 2 else float("nan")
    if rank == 0:
        print(f"world {world} {mib} MiB unroll {unroll} blocks {nb}: two-shot {t:.1f} us, one-shot {t1:.1f} us per all-reduce (max over ranks)", flush=True)
tn = timed(lambda: [dist.all_reduce(p) for p in plain])
if rank == 0:
    print(f"world {world} {mib} MiB: NCCL {tn:.1f} us per all-reduce", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
