"""Development aid (torchrun, >= 2 GPUs, HG_EXTRA_NVCC_FLAGS=-DHG_OPROJ_TRACE at build and run time): %globaltimer stamps of
the stages of one fused o_proj + all-reduce launch, per CTA, relative to the earliest CTA entry."""
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
M, N, KF, NL = 1024, 4096, 4096, 8
K = KF // world
dt = torch.bfloat16
ar = MultimemAllReduce(NL * (M * N * 2 + 256) + 4096, dev)
xs = [torch.randn(M, K, device=dev).to(dt) for _ in range(NL)]
ws = [(torch.randn(N, K, device=dev) / KF**0.5).to(dt) for _ in range(NL)]
bufs = [ar.buffer((M, N), dt) for _ in range(NL)]
lib = _lib.load()


def fused():
    for x, w, b in zip(xs, ws, bufs):
        ar.linear_all_reduce_(x, w, b)


fused()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    fused()
for _ in range(5):
    gr.replay()
torch.cuda.synchronize()
dist.barrier()
gr.replay()
torch.cuda.synchronize()
n_cta = torch.cuda.get_device_properties(dev).multi_processor_count
buf = (ctypes.c_longlong * (160 * 16))()
lib.hg_debug_oproj_trace(buf, 160 * 16)
rows = [[buf[c * 16 + s] for s in range(16)] for c in range(n_cta)]
t0 = min(r[0] for r in rows)
if rank == 0:
    names = {0: "entry", 1: "dep-wait passed", 2: "first accumulator full", 8: "first tile signalled", 10: "MMA warp done", 3: "last tile signalled",
             4: "first slice ready", 9: "first slice multicast", 5: "reduce warp 0 done", 6: "fence.sys done"}
    print(f"world {world} [{M},{N}] k={K}: stage times in us after the first CTA's entry (min / median / max over CTAs)", flush=True)
    for s, nm in names.items():
        v = sorted((r[s] - t0) / 1e3 for r in rows if r[s] >= t0)
        if v:
            print(f"  {nm:24s} {v[0]:7.2f} {v[len(v) // 2]:7.2f} {v[-1]:7.2f}   ({len(v)} CTAs)", flush=True)
    last = max(rows, key=lambda r: r[7])
    print(f"  last CTA: end barrier passed {(last[7] - t0) / 1e3:.2f}", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
