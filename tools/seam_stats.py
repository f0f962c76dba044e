"""Where one stages.main(float64 numpy) call spends its time on the host side: per configuration of the host
transport (split of the conversion work, chunk size, ring length, download route) the wall time per call on
rotating inputs and, with MGB_HOST_STATS, the transport's own account of two uploads.
GPU box only:  python tools/seam_stats.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
import numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "oracle"))
import port, matchering_b200 as mg
from matchering_b200 import stages
from matchering_b200.sharding import bind_host_thread_near_gpu
bind_host_thread_near_gpu(0)
n = 44100 * 180
cfg = mg.Config()
ts = [port.synth_target(n, k).astype(np.float64) for k in range(3)]
rs = [port.synth_reference(n, 100 + k).astype(np.float64) for k in range(3)]
for k in range(4):
    stages.main(ts[k %% 3], rs[k %% 3], cfg)
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 15
for k in range(K):
    out = stages.main(ts[k %% 3], rs[k %% 3], cfg)[0]
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / K * 1e3
print(f"{ms:.2f} ms per call -> {180e3 / ms:.0f}x real-time")
''' % (ROOT, ROOT)

CONFIGS = [  # (split, chunk samples, ring, download through the ring)
    ("slice", 65536, 16, 0), ("chunk", 65536, 32, 0), ("slice", 262144, 8, 0), ("slice", 1048576, 6, 0)]
CONFIGS_OLD = [
    ("slice", 65536, 16, 0), ("slice", 65536, 16, 1), ("chunk", 65536, 16, 0), ("chunk", 65536, 32, 0),
    ("chunk", 32768, 32, 0), ("chunk", 32768, 48, 0), ("slice", 131072, 8, 0), ("slice", 32768, 32, 0),
    ("chunk", 65536, 24, 1),
]
threads = os.environ.get("MGB_HOST_THREADS", "12")
for split, chunk, ring, dring in CONFIGS:
    env = dict(os.environ, MGB_HOST_THREADS=threads, MGB_HOST_SPLIT=split, MGB_HOST_CHUNK=str(chunk), MGB_HOST_RING=str(ring),
               MGB_DOWNLOAD_RING=str(dring))
    env.pop("MGB_HOST_STATS", None)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"threads={threads} split={split} chunk={chunk} ring={ring} ring_download={dring}: "
          f"{r.stdout.strip() or r.stderr.strip()[-300:]}", flush=True)
    if True:
        r = subprocess.run([sys.executable, "-c", CHILD], env=dict(env, MGB_HOST_STATS="1"), capture_output=True, text=True)
        lines = [l for l in r.stderr.splitlines() if l.startswith("[mgb upload]")]
        print("\n".join("    " + l for l in lines[-4:]), flush=True)
