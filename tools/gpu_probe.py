"""Per-stage device timings of one mastering job (CUDA events), for quick looks on the GPU box."""
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
from matchering_b200.engine import TrackSession, get_plan, to_device_f32  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
    n = int(44100 * seconds)
    cfg = mg.Config()
    t0 = time.time()
    plan = get_plan(cfg)
    torch.cuda.synchronize()
    print(f"plan build + upload {time.time() - t0:.3f} s")
    t = to_device_f32(port.synth_target(n, 0), plan.device)
    r = to_device_f32(port.synth_reference(n, 1), plan.device)
    s = TrackSession(plan, n, n)
    for tma in (1, 0):
        plan.lib.mgb_set_option(b"tma", tma)
        for it in range(3):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
            s.match_levels(t, r); ev[1].record()
            s.match_frequencies(t); ev[2].record()
            s.correct_levels(); ev[3].record()
            out = s.finalize(True, False, False); ev[4].record()
            torch.cuda.synchronize()
            ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
        total = sum(ms)
        print(f"tma={tma} {seconds:.0f}s track: levels {ms[0]:.3f} freq {ms[1]:.3f} correct {ms[2]:.3f} "
              f"finalize {ms[3]:.3f} total {total:.3f} ms -> {seconds / (total * 1e-3):.0f}x real-time")
    st = s.read_state()
    print("engaged", st.limiter_engaged, "gain", st.gain, "peak", st.result_peak)


if __name__ == "__main__":
    main()
