"""Where the time of stages.main(float64 numpy) goes, and how it moves with the host transport's geometry:
upload (worker threads narrow into the pinned ring, link copies), kernels, download; threads x chunk sweep.
Run on the GPU box:  python tools/seam_sweep.py [seconds]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import port  # noqa: E402
import matchering_b200 as mg  # noqa: E402
from matchering_b200 import _native, stages  # noqa: E402
from matchering_b200.engine import HostIO, get_plan, host_session  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
n = int(44100 * seconds)
cfg = mg.Config()
plan = get_plan(cfg)
lib = plan.lib
t64 = port.synth_target(n, 0).astype(np.float64)
r64 = port.synth_reference(n, 1).astype(np.float64)
sess = host_session(plan, n, n)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
pool_out = HostIO.get().pool.array((n, 2), np.float64)


def wall(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print(f"track: {seconds:.0f} s, {n} frames; host: {os.cpu_count()} hardware threads")
for threads in (4, 8, 16, 24, 32, 48, 64):
    for chunk in (1 << 18, 1 << 20, 1 << 22):
        h = C.c_void_p()
        _native.check(lib, lib.mgb_host_io_create(threads, chunk, 6, C.byref(h)))
        up = wall(lambda: (_native.check(lib, lib.mgb_host_upload(h, t64.ctypes.data, 8, sess.d_target.data_ptr(), 2 * n, stream)),
                           torch.cuda.synchronize()))
        down_pinned = wall(lambda: _native.check(lib, lib.mgb_host_download(h, sess.d_out.data_ptr(), pool_out.ctypes.data, 8, 2 * n,
                                                                             sess.d_wide.data_ptr(), stream)))
        pageable = np.empty((n, 2), dtype=np.float64)
        down_page = wall(lambda: _native.check(lib, lib.mgb_host_download(h, sess.d_out.data_ptr(), pageable.ctypes.data, 8, 2 * n,
                                                                           None, stream)))
        print(f"threads {threads:3d} chunk {chunk >> 10:5d} Ki samples: upload f64->f32 {up:6.2f} ms ({2 * n * 8 / up / 1e6:6.1f} GB/s read), "
              f"download to pinned f64 {down_pinned:6.2f} ms, to pageable f64 {down_page:6.2f} ms")
        lib.mgb_host_io_destroy(h)

for threads in (8, 16, 32, 64):
    os.environ["MGB_HOST_THREADS"] = str(threads)
    HostIO._instance = None
    ms = wall(lambda: stages.main(t64, r64, cfg), reps=10)
    print(f"stages.main(float64 numpy) with {threads} worker threads: {ms:.2f} ms -> {seconds / ms * 1e3:.0f}x real-time")
