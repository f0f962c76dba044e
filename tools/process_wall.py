"""Wall-clock of mg.process on WAV files (file read, device mastering, file write)."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import port  # noqa: E402
import matchering_b200 as mg  # noqa: E402
from matchering_b200 import wavio  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
n = int(44100 * seconds)
d = tempfile.mkdtemp()
for subtype in ("PCM_16", "PCM_24", "FLOAT"):
    wavio.write(os.path.join(d, "t.wav"), port.synth_target(n, 0), 44100, subtype)
    wavio.write(os.path.join(d, "r.wav"), port.synth_reference(n, 1), 44100, subtype)
    for it in range(3):
        t0 = time.perf_counter()
        mg.process(os.path.join(d, "t.wav"), os.path.join(d, "r.wav"), [mg.pcm16(os.path.join(d, "o16.wav"))])
        dt = time.perf_counter() - t0
    print(f"{subtype}: mg.process on a {seconds:.0f}-s track: {dt * 1e3:.1f} ms wall ({seconds / dt:.0f}x real-time), files included")
