"""cProfile of mg.process on 16-bit WAV files in /dev/shm (where does the host time go?).  GPU box only."""
import cProfile
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import port  # noqa: E402
import matchering_b200 as mg  # noqa: E402
from matchering_b200 import wavio  # noqa: E402

n = 44100 * 180
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
wavio.write(os.path.join(d, "t.wav"), port.synth_target(n, 0), 44100, "PCM_16")
wavio.write(os.path.join(d, "r.wav"), port.synth_reference(n, 1), 44100, "PCM_16")
run = lambda: mg.process(os.path.join(d, "t.wav"), os.path.join(d, "r.wav"), [mg.pcm16(os.path.join(d, "o.wav"))])
for _ in range(3):
    t0 = time.perf_counter()
    run()
    print(f"mg.process: {(time.perf_counter() - t0) * 1e3:.1f} ms")
prof = cProfile.Profile()
prof.enable()
run()
prof.disable()
pstats.Stats(prof).sort_stats("cumulative").print_stats(30)
