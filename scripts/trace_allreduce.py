"""Development aid (torchrun, >= 2 GPUs, library built with HG_EXTRA_NVCC_FLAGS=-DHG_AR_TRACE and the same variable set
at run time): %globaltimer stamps of the stages of one hg_allreduce_multimem call, per CTA, relative to the earliest
CTA entry: where the microseconds of an 8 MiB all-reduce go (start barrier / switch reductions + multicast stores /
fence.sys / end barrier)."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hydragen_b200 import _lib  # noqa: E402
from hydragen_b200.collectives import MultimemAllReduce  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
B, HID, NL = 1024, 4096, 8
ar = MultimemAllReduce(NL * B * HID * 2 + 4096, dev)
bufs = [ar.buffer((B, HID), torch.bfloat16).zero_() for _ in range(NL)]
lib = _lib.load()
for nb in [int(x) for x in os.environ.get("AR_BLOCKS", "16,64").split(",")]:
    ar.n_blocks = nb
    gr = torch.cuda.CUDAGraph()
    for b in bufs:
        ar.all_reduce_(b)
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        for b in bufs:
            ar.all_reduce_(b)
    for _ in range(5):
        gr.replay()
    torch.cuda.synchronize()
    dist.barrier()
    gr.replay()
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (256 * 8))()
    lib.hg_debug_ar_trace(buf, 256 * 8)
    rows = [[buf[c * 8 + s] for s in range(8)] for c in range(nb)]
    t0 = min(r[0] for r in rows)
    if rank == 0:
        names = ["entry", "dep-wait", "start-bar", "data", "fence", "signalled", "end-bar"]
        print(f"world {world} blocks {nb}: stage completion times in us after the first CTA's entry (min / max over CTAs)", flush=True)
        for s in range(5):
            v = [(r[s] - t0) / 1e3 for r in rows]
            print(f"  {names[s]:10s} {min(v):7.2f} {max(v):7.2f}", flush=True)
        last = max(rows, key=lambda r: r[6])  # only the last CTA of the most recent call refreshed stamp 6
        print(f"  last CTA: signalled {(last[5] - t0) / 1e3:.2f}, end barrier passed {(last[6] - t0) / 1e3:.2f}", flush=True)
    del gr
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
