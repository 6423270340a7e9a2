"""Development aid (torchrun, >= 2 GPUs): graph-timed latency of hg_allreduce_multimem (two-shot in place, one-shot out
of place) at the decode-step message size; AR_BLOCKS = CTAs.  (r01n used it with temporary debug switches in the kernel
to split the 45.8 us of an 8 MiB all-reduce on 2 GPUs: DESIGN.md section 5.)"""
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
nb = int(os.environ.get("AR_BLOCKS", "128"))
ar = MultimemAllReduce(NL * B * HID * 2 + 4096, dev, n_blocks=nb)
assert ar.available
bufs = [ar.buffer((B, HID), torch.bfloat16).zero_() for _ in range(NL)]


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


t = timed(lambda: [ar.all_reduce_(b) for b in bufs])
outs = [torch.empty(B, HID, device=dev, dtype=torch.bfloat16) for _ in range(NL)]
t1 = timed(lambda: [ar.all_reduce(b, o) for b, o in zip(bufs, outs)])
if rank == 0:
    print(f"world {world} blocks {nb}: "
          f"two-shot {t:.1f} us, one-shot {t1:.1f} us per 8 MiB all-reduce", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
