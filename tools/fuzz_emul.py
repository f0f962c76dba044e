"""Randomised differential run: the kernels on the CPU emulator against the oracle port over random Configs, lengths
and levels (no GPU).   python tools/fuzz_emul.py [seed] [cases]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul")):
    if p not in sys.path:
        sys.path.insert(0, p)
import port  # noqa: E402  (the oracle: the checker)
from emul_harness import run_pipeline  # noqa: E402
from matchering_b200 import plan as plan_mod  # noqa: E402


def random_case(rng) -> dict:
    F = int(rng.choice([512, 1024, 2048, 4096, 4096, 8192, 16384]))
    sr = int(rng.choice([8000, 22050, 44100, 48000, 96000, 176400, 192000]))
    lo = F + 1 + int(rng.integers(0, 50))
    hi = max(lo + 1, min(int(sr * 3.0), 200000))
    min_piece_s = (F + 2) / sr
    return dict(
        fft_size=F, sample_rate=sr, n=int(rng.integers(lo, hi)), nr=int(rng.integers(lo, hi)),
        max_piece_size=float(rng.uniform(min_piece_s * 1.01, max(min_piece_s * 1.5, 2.0))),
        steps=int(rng.integers(0, 6)), threshold=float(rng.choice([0.998138, 0.9, 0.5])),
        attack=float(rng.choice([1.0, 0.5, 2.0])), hold=float(rng.choice([1.0, 0.5, 3.0])),
        release=float(rng.choice([3000.0, 500.0, 6000.0])), hold_order=int(rng.choice([1, 1, 2])),
        release_order=int(rng.choice([1, 1, 2])), lowess_it=int(rng.choice([0, 0, 0, 1, 2])),
        loud_t=float(rng.choice([1.0, 0.05, 1e-4, 3.0])), loud_r=float(rng.choice([1.0, 0.02, 2.5])),
        seed_t=int(rng.integers(0, 1000)), seed_r=int(rng.integers(0, 1000)), mono=bool(rng.random() < 0.15),
        silent_start=bool(rng.random() < 0.1))


def run_case(c: dict):
    """-> ("ok" | "skipped" | "failed", error relative to max(1, peak of the reference result), description)"""
    desc = ", ".join(f"{k}={v}" for k, v in c.items())
    lim = port.OracleLimiterConfig(attack=c["attack"], hold=c["hold"], release=c["release"],
                                   hold_filter_order=c["hold_order"], release_filter_order=c["release_order"])
    cfg = port.OracleConfig(internal_sample_rate=c["sample_rate"], fft_size=c["fft_size"], max_piece_size=c["max_piece_size"],
                            rms_correction_steps=c["steps"], threshold=c["threshold"], limiter=lim, lowess_it=c["lowess_it"])
    try:
        plan_mod.build_tables(cfg)
    except plan_mod.UnsupportedConfig as e:
        return "skipped", 0.0, f"unsupported ({e}): {desc}"
    t = (port.synth_target(c["n"], c["seed_t"]) * c["loud_t"]).astype(np.float32)
    r = (port.synth_reference(c["nr"], c["seed_r"]) * c["loud_r"]).astype(np.float32)
    if c["mono"]:
        t[:, 1] = t[:, 0]
    if c["silent_start"]:
        t[: c["n"] // 3] = 0.0
    # pieces shorter than fft_size: scipy's STFT changes its own frame length there, the package rejects the layout
    for frames in (c["n"], c["nr"]):
        divisions = int(frames / (c["max_piece_size"] * c["sample_rate"])) + 1
        if int(frames / divisions) < c["fft_size"]:
            return "skipped", 0.0, f"piece shorter than fft_size: {desc}"
    want = port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True)
    try:
        outs, _, _, _, _ = run_pipeline(cfg, t, r)
    except Exception as e:  # noqa: BLE001
        return "failed", float("inf"), f"{type(e).__name__}: {e}: {desc}"
    err = 0.0
    for a, b in zip(outs, want):
        err = max(err, float(np.abs(np.asarray(a, dtype=np.float64) - b).max()) / max(1.0, float(np.abs(b).max())))
    return "ok", err, desc


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rng = np.random.default_rng(seed)
    worst = 0.0
    for i in range(cases):
        t0 = time.time()
        outcome, err, desc = run_case(random_case(rng))
        worst = max(worst, err if outcome == "ok" else 0.0)
        flag = "  <<<<<<<< MISMATCH" if outcome == "failed" or err > 1e-5 else ""
        print(f"case {i}: {outcome} {err:.2e} ({time.time() - t0:.1f} s) :: {desc}{flag}", flush=True)
    print("worst", worst)
