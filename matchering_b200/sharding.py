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
