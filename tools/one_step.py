"""Two device-resident mastering steps of config 2 and nothing else: the command ncu wraps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import port  # noqa: E402
import matchering_b200 as mg  # noqa: E402
from matchering_b200.engine import TrackSession, get_plan, to_device_f32  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(44100 * seconds)
plan = get_plan(mg.Config())
t = to_device_f32(port.synth_target(n, 0), plan.device)
r = to_device_f32(port.synth_reference(n, 1), plan.device)
s = TrackSession(plan, n, n)
for _ in range(steps):
    s.match_levels(t, r)
    s.match_frequencies(t)
    s.correct_levels()
    s.finalize(True, False, False)
torch.cuda.synchronize()
print("done")
