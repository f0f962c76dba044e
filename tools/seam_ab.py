"""A/B of the host transport's geometry INSIDE one process: every configuration gets its own mgb_host_io, the
calls of all configurations alternate round-robin on rotating inputs (so drift, throttling and cache state hit
all of them alike) and the median / minimum wall time of mgb_stages_main_host per configuration is printed.
GPU box only:  python tools/seam_ab.py [rounds]"""
import ctypes as C
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import port  # noqa: E402
import matchering_b200 as mg  # noqa: E402
from matchering_b200 import _native  # noqa: E402
from matchering_b200.engine import HostIO, get_plan, host_session  # noqa: E402
from matchering_b200.sharding import bind_host_thread_near_gpu  # noqa: E402

bind_host_thread_near_gpu(0)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 25
n = 44100 * 180
cfg = mg.Config()
plan = get_plan(cfg)
lib = plan.lib
ts = [port.synth_target(n, k).astype(np.float64) for k in range(3)]
rs = [port.synth_reference(n, 100 + k).astype(np.float64) for k in range(3)]
sess = host_session(plan, n, n)
out = HostIO.get().pool.array((n, 2), np.float64)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
state = _native.TrackState()

# (label, threads, chunk samples, ring, streaming stores (0 plain, 1 / 2 = 256 / 512 bit), download route (0 one DMA of
#  device-widened float64, 2 float32 chunks through the ring), whole chunks per worker, prefetch bytes)
CONFIGS = [
    ("t13 64K x16 plain  dma ", 13, 1 << 16, 16, 0, 0, 0, 0),
    ("t13 512K x8 nt512  dma ", 13, 1 << 19, 8, 2, 0, 0, 0),
    ("t15 512K x8 nt512  dma ", 15, 1 << 19, 8, 2, 0, 0, 0),
    ("t13 512K x8 nt512  ring pf8k", 13, 1 << 19, 8, 2, 2, 0, 8192),
    ("t13 1M x6   nt512  ring pf8k", 13, 1 << 20, 6, 2, 2, 0, 8192),
    ("t13 1M x6   nt512  ring", 13, 1 << 20, 6, 2, 2, 0, 0),
    ("t13 2M x4   nt512  ring pf8k", 13, 1 << 21, 4, 2, 2, 0, 8192),
    ("t15 1M x6   nt512  ring pf8k", 15, 1 << 20, 6, 2, 2, 0, 8192),
]
if len(sys.argv) > 2:
    CONFIGS = [c for c in CONFIGS if any(key in c[0] for key in sys.argv[2:])]
handles = []
for label, threads, chunk, ring, nt, dring, whole, pf in CONFIGS:
    h = C.c_void_p()
    _native.check(lib, lib.mgb_host_io_create(threads, chunk, ring, C.byref(h)))
    handles.append(h)


def call(i, k):
    label, threads, chunk, ring, nt, dring, whole, pf = CONFIGS[i]
    for name, value in (("host_streaming_stores", nt), ("host_download_ring", dring), ("host_split_chunks", whole), ("host_prefetch", pf)):
        _native.check(lib, lib.mgb_set_option(name.encode(), value))
    t0 = time.perf_counter()
    _native.check(lib, lib.mgb_stages_main_host(handles[i], C.byref(plan.struct), C.byref(sess.layout), ts[k % 3].ctypes.data,
                                                 rs[k % 3].ctypes.data, 8, out.ctypes.data, None, None, 8, C.byref(sess.host_buffers),
                                                 C.byref(state), stream))
    return (time.perf_counter() - t0) * 1e3


times = [[] for _ in CONFIGS]
k = 0
for r in range(rounds + 2):
    for i in range(len(CONFIGS)):
        ms = call(i, k)
        k += 1
        if r >= 2:
            times[i].append(ms)
print(f"stages.main host seam, 180-s track, {rounds} alternating rounds; ms per call: median / min / max")
for (label, *_), t in zip(CONFIGS, times):
    print(f"  {label}: {statistics.median(t):6.2f} / {min(t):6.2f} / {max(t):6.2f}")
for h in handles:
    lib.mgb_host_io_destroy(h)
