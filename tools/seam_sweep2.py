"""stages.main(float64 numpy) on ROTATING inputs (3 tracks: 760 MB of sources, no cache help), across worker
threads, chunk sizes, download route and core binding.  GPU box only:  python tools/seam_sweep2.py"""
import itertools
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
if os.environ.get("SWEEP_BIND") == "1":
    from matchering_b200.sharding import bind_host_thread_near_gpu
    bound = bind_host_thread_near_gpu(0)
else:
    bound = None
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
print(f"bound={bound} threads={os.environ.get('MGB_HOST_THREADS')} chunk={os.environ.get('MGB_HOST_CHUNK')} "
      f"ring_download={os.environ.get('MGB_DOWNLOAD_RING', '0')}: {ms:.2f} ms per call -> {180e3 / ms:.0f}x real-time")
''' % (ROOT, ROOT)
for bind, threads, chunk, ring in itertools.product(("0", "1"), ("8", "16", "32"), ("65536", "262144"), ("0", "1")):
    env = dict(os.environ, SWEEP_BIND=bind, MGB_HOST_THREADS=threads, MGB_HOST_CHUNK=chunk, MGB_DOWNLOAD_RING=ring)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
