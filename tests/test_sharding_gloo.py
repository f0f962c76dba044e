"""The multi-GPU story is 'tracks shard one-per-rank, no data-path collective, one gather of the
timings' (SURVEY.md 8e).  This runs that host logic with world_size 2 on the gloo backend."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from matchering_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tracks = list(range(7))
    mine = sharding.tracks_for_rank(len(tracks), rank, world)
    elapsed_ms = 10.0 * (rank + 1)          # pretend rank 1 is the slow one
    frames = sum(1000 * (t + 1) for t in mine)
    total_frames, max_ms = sharding.gather_throughput(frames, elapsed_ms, device=None)
    out.put((rank, mine, total_frames, max_ms))
    dist.destroy_process_group()


def test_round_robin_sharding_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1] == [0, 2, 4, 6] and results[1][1] == [1, 3, 5]
    want_frames = sum(1000 * (t + 1) for t in range(7))
    for _, _, total, max_ms in results:
        assert total == want_frames and max_ms == 20.0


def test_single_process_is_identity():
    sys.path.insert(0, ROOT)
    from matchering_b200 import sharding
    assert sharding.tracks_for_rank(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.gather_throughput(123, 4.5, device=None) == (123, 4.5)


def test_numa_binding_is_a_no_op_without_a_gpu():
    import os
    import torch
    from matchering_b200.sharding import bind_host_thread_near_gpu
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    before = os.sched_getaffinity(0)
    assert bind_host_thread_near_gpu(0) is None and os.sched_getaffinity(0) == before
