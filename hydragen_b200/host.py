"""Host-buffer entry point of the decode hot path: the caller's q / k_new / v_new live in pinned HOST memory and
the attention output is wanted back on the host (an engine whose projections run elsewhere, or a measurement
through the boundary with the copies included).  Per layer the work is

    H2D(q, k_new, v_new)  ->  prefix launch(es) + fused append/suffix/combine launch  ->  D2H(out)

and a decode step is that for every layer.  The three stages run on three streams and are chained per layer with
events, so layer i+1's upload and layer i-1's download overlap layer i's kernels (PCIe is full duplex); a layer's
staging buffers are only rewritten after the previous step's kernels of that layer have finished.  torch supplies
streams, events and pinned memory; the kernels are the C-ABI library's.

``capture()`` records one whole step -- every layer's copies and kernels on the three streams, with their
dependencies -- into ONE CUDA graph (the host and staging buffers are static), so a step costs one graph launch
instead of ~10 Python-issued stream operations per layer: the step is then bound by PCIe, not by launch overhead
(round-1 review: the eager form scaled 2.5x from 1 to 8 GPUs where the 8 independent PCIe links allow ~8x).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from .attention import hydragen_attention_decode


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' (the sysfs cpulist format) -> [0, 1, 2, 3, 8, 10, 11]."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_process_to_gpu_numa(device_index: int, sysfs_root: str = "/sys/bus/pci/devices") -> Optional[dict]:
    """Restrict this process to the CPUs that are local to GPU ``device_index`` (its PCIe root's NUMA node, read from sysfs), so
    that the pinned host buffers allocated afterwards -- first touch -- sit on the socket the GPU hangs off.  One process per GPU
    (torchrun); call it before allocating pinned memory.  Returns what was done, or None when there is nothing to do or anything
    about the platform is unexpected (no sysfs entry, no NUMA information, CPUs outside this process's cpuset): it never raises.
    ``HYDRAGEN_B200_BIND_NUMA=0`` disables it."""
    import os

    if os.environ.get("HYDRAGEN_B200_BIND_NUMA", "1") == "0":
        return None
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(os.path.join(sysfs_root, bdf, "local_cpulist")) as f:
            local = set(parse_cpulist(f.read()))
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if not target or target == allowed:
            return None
        os.sched_setaffinity(0, target)
        node = None
        try:
            with open(os.path.join(sysfs_root, bdf, "numa_node")) as f:
                node = int(f.read().strip())
        except Exception:
            pass
        return {"pci": bdf, "numa_node": node, "cpus": len(target), "of": len(allowed)}
    except Exception:
        return None


@dataclass
class HostDecodeLayer:
    """One layer's operands.  ``*_host`` are pinned host tensors, ``*_dev`` the device staging buffers of the
    same shapes, the caches and shared K/V are resident on the device."""

    q_host: Tensor
    k_host: Tensor
    v_host: Tensor
    out_host: Tensor
    q_dev: Tensor
    k_dev: Tensor
    v_dev: Tensor
    k_cache: Tensor
    v_cache: Tensor
    shared_ks: List[Tensor]
    shared_vs: List[Tensor]
    shared_cu_seq_lens: Optional[List[Optional[Tensor]]] = None
    shared_max_seq_lens: Optional[List[Optional[int]]] = None
    use_varlens: Optional[List[bool]] = None
    out_dev: Optional[Tensor] = field(default=None, repr=False)


class HostDecodePipeline:
    def __init__(self, device: torch.device):
        self.device = device
        self.h2d = torch.cuda.Stream(device)
        self.d2h = torch.cuda.Stream(device)
        self._done: dict = {}  # layer index -> event: kernels of the previous step finished (staging reusable)

    def step(self, layers: Sequence[HostDecodeLayer], positions: Tensor, after_layer=None) -> None:
        """One decode step over ``layers``; ``positions [b]`` (device) = row of the new token in the unique
        caches.  ``after_layer(i)`` (optional) is called on the compute stream after layer i's kernels (the
        tensor-parallel all-reduce goes there).  Returns once everything is enqueued; ``out_host`` is valid
        after ``synchronize()``."""
        compute = torch.cuda.current_stream(self.device)
        for i, ly in enumerate(layers):
            with torch.cuda.stream(self.h2d):
                ev = self._done.get(i)
                if ev is not None:
                    self.h2d.wait_event(ev)
                ly.q_dev.copy_(ly.q_host, non_blocking=True)
                ly.k_dev.copy_(ly.k_host, non_blocking=True)
                ly.v_dev.copy_(ly.v_host, non_blocking=True)
                up = torch.cuda.Event()
                up.record(self.h2d)
            compute.wait_event(up)
            ly.out_dev = hydragen_attention_decode(ly.q_dev, ly.k_dev, ly.v_dev, positions, ly.k_cache, ly.v_cache, ly.shared_ks, ly.shared_vs,
                                                   ly.shared_cu_seq_lens, ly.shared_max_seq_lens, ly.use_varlens)
            if after_layer is not None:
                after_layer(i)
            done = torch.cuda.Event()
            done.record(compute)
            self._done[i] = done
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(done)
                ly.out_host.copy_(ly.out_dev, non_blocking=True)
                ly.out_dev.record_stream(self.d2h)

    def capture(self, layers: Sequence[HostDecodeLayer], positions: Tensor, after_layer=None) -> "torch.cuda.CUDAGraph":
        """One decode step over ``layers`` as a CUDA graph: the upload / kernel / download streams fork from the
        capturing stream and join it again, so replaying the graph on a stream orders whole steps on that stream.
        Run ``step`` once eagerly beforehand (library set-up, workspace, collective arenas must exist)."""
        graph = torch.cuda.CUDAGraph()
        self.synchronize()
        self._done = {}
        with torch.cuda.graph(graph):
            cap = torch.cuda.current_stream(self.device)
            self.h2d.wait_stream(cap)
            self.d2h.wait_stream(cap)
            for i, ly in enumerate(layers):
                with torch.cuda.stream(self.h2d):
                    ly.q_dev.copy_(ly.q_host, non_blocking=True)
                    ly.k_dev.copy_(ly.k_host, non_blocking=True)
                    ly.v_dev.copy_(ly.v_host, non_blocking=True)
                    up = torch.cuda.Event()
                    up.record(self.h2d)
                cap.wait_event(up)
                ly.out_dev = hydragen_attention_decode(ly.q_dev, ly.k_dev, ly.v_dev, positions, ly.k_cache, ly.v_cache, ly.shared_ks, ly.shared_vs,
                                                       ly.shared_cu_seq_lens, ly.shared_max_seq_lens, ly.use_varlens)
                if after_layer is not None:
                    after_layer(i)
                done = torch.cuda.Event()
                done.record(cap)
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(done)
                    ly.out_host.copy_(ly.out_dev, non_blocking=True)
            cap.wait_stream(self.h2d)
            cap.wait_stream(self.d2h)
        return graph

    def synchronize(self) -> None:
        self.h2d.synchronize()
        self.d2h.synchronize()
        torch.cuda.current_stream(self.device).synchronize()
