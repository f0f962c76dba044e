"""Multi-GPU plumbing: tracks are independent, so they shard round-robin one process per GPU with
no exchange on the data path; the only collective gathers (frames, elapsed) for the aggregate
throughput (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def tracks_for_rank(n_tracks: int, rank: int, world: int) -> list:
    return list(range(rank, n_tracks, world))


def gather_throughput(frames: int, elapsed_ms: float, device=None):
    """-> (total frames over all ranks, max elapsed over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return frames, elapsed_ms
    mine = torch.tensor([float(frames), float(elapsed_ms)], dtype=torch.float64, device=device)
    gathered = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, mine)
    return int(sum(float(g[0]) for g in gathered)), max(float(g[1]) for g in gathered)


def bind_host_thread_near_gpu(device_index: int):
    """One process per GPU moves ~190 MB per track over PCIe from pinned host memory: keep the
    process (and, by first touch, its pinned buffers) on the CPU cores NVML reports as local to the
    GPU, so eight ranks do not drag each other's copies across the socket interconnect.
    -> number of cores bound to, or None when NVML cannot tell (nothing is changed then)."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cores = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return len(allowed)
    except Exception:  # no NVML, no permission, unknown topology: run unbound
        return None
