"""Where do the pinned result blocks of stages.main(numpy) go when the arrays are dropped?  (GPU box)"""
import gc
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import port  # noqa: E402
import matchering_b200 as mg  # noqa: E402
from matchering_b200 import stages  # noqa: E402
from matchering_b200.engine import HostIO  # noqa: E402

cfg = mg.Config(max_piece_size=2.0)
t, r = port.synth_target(150000, 71).astype(np.float64), port.synth_reference(140000, 72).astype(np.float64)
pool = None


def show(tag):
    global pool
    pool = HostIO.get().pool
    print(f"{tag}: cached {pool.cached}, free lists { {k: len(v) for k, v in pool.free.items()} }, keep_bytes {pool.keep_bytes}", flush=True)


a = stages.main(t, r, cfg)[0]
show("after a")
print("a base chain:", type(a.base), sys.getrefcount(a.base), flush=True)
b = stages.main(t * 0.5, r, cfg)[0]
show("after b")
addr = a.ctypes.data
view = a[::2]
del a
gc.collect()
show("after del a (view alive)")
c = stages.main(t, r, cfg)[0]
show("after c")
print("c is new block:", c.ctypes.data != addr, flush=True)
del view, c
gc.collect()
show("after del view, c")
d = stages.main(t, r, cfg)[0]
show("after d")
del d, b
gc.collect()
show("after del d, b")
