"""ORACLE (test infrastructure only) -- golden vectors for BASELINE configs 3 and 5 AT FULL SIZE.

Runs the UNMODIFIED reference (/root/reference through oracle/ref_shims.py) in the build container on
the SURVEY.md section 8(d) recipes

  C5   limiter.limit() on one hour of 44.1 kHz stereo (158.76 M frames; ~14 GB resident, ~2 min)
  C3   stages.main() on ten minutes of 96 kHz stereo (57.6 M frames; ~1 min)

and keeps DECIMATED outputs, since the full arrays (1.3 GB / 0.5 GB as float32) cannot be committed:
every `every`-th frame (a prime stride, so every phase of the 4608-sample limiter chunks, the
4096-sample convolution frames and the pieces is visited), `windows` contiguous stretches spread over
the buffer (placed across limiter chunk / convolution frame boundaries deep inside it), 64 block sums
of the output and of its square, the peak and where it sits.  The inputs are NOT stored: both sides
regenerate them from oracle/port.py's seeded recipes; a float64 sum of the float32 input is kept so a
test can tell "different input" from "different output".

    python oracle/make_golden_full.py [c5] [c3]          (build container only)
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import port  # noqa: E402
from ref_shims import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
EVERY = 997
N_WINDOWS = 12
WINDOW = 2048
N_BLOCKS = 64


def window_starts(n: int, period: int) -> np.ndarray:
    """N_WINDOWS stretches of WINDOW frames, each straddling a multiple of `period` (a limiter chunk or
    a convolution frame boundary), spread from the first tenth of the buffer to its very end."""
    anchors = np.linspace(0.1 * n, n - WINDOW, N_WINDOWS)
    starts = (np.round(anchors / period) * period - WINDOW // 2).astype(np.int64)
    starts = np.clip(starts, 0, n - WINDOW)
    starts[-1] = n - WINDOW  # the tail of the buffer itself (filtfilt's end extension, last partial chunk)
    return starts


def decimate(y: np.ndarray, period: int) -> dict:
    n = y.shape[0]
    starts = window_starts(n, period)
    edges = np.linspace(0, n, N_BLOCKS + 1).astype(np.int64)
    mono = np.abs(y).max(axis=1)
    peak_at = int(np.argmax(mono))
    return dict(
        frames=n, every=EVERY, rows=y[::EVERY].astype(np.float64),
        window=WINDOW, window_starts=starts, windows=np.stack([y[s:s + WINDOW] for s in starts]).astype(np.float64),
        block_edges=edges,
        block_sum=np.array([y[a:b].sum(axis=0) for a, b in zip(edges[:-1], edges[1:])]),
        block_sumsq=np.array([np.einsum("ij,ij->j", y[a:b], y[a:b]) for a, b in zip(edges[:-1], edges[1:])]),
        peak=float(mono[peak_at]), peak_at=peak_at)


def make_c5(matchering):
    from matchering import Config
    from matchering.limiter import limit
    n = 44100 * 3600
    t0 = time.time()
    x = port.synth_limiter_input(n, seed=0)
    x64 = x.astype(np.float64)
    input_sum = float(x64.sum())
    print(f"c5: input ready ({time.time() - t0:.0f} s)", flush=True)
    t0 = time.time()
    y = limit(x64, Config())
    print(f"c5: reference limit() took {time.time() - t0:.0f} s", flush=True)
    d = decimate(y, 4608)
    np.savez_compressed(os.path.join(OUT, "c5_limiter_hour.npz"), input_sum=input_sum, seed=0, sample_rate=44100, **d)


def make_c3(matchering):
    from matchering import Config, stages
    n = 96000 * 600
    t0 = time.time()
    t = port.synth_target(n, 0)
    r = port.synth_reference(n, 1)
    t64, r64 = t.astype(np.float64), r.astype(np.float64)
    input_sum = float(t64.sum() + r64.sum())
    del t, r
    print(f"c3: inputs ready ({time.time() - t0:.0f} s)", flush=True)
    t0 = time.time()
    cfg = Config(internal_sample_rate=96000)
    limited, plain, _ = stages.main(t64, r64, cfg, need_default=True, need_no_limiter=True)
    print(f"c3: reference stages.main() took {time.time() - t0:.0f} s", flush=True)
    d_lim = decimate(limited, 4096)
    d_plain = decimate(plain, 4096)
    out = {"limited_" + k: v for k, v in d_lim.items()}
    out.update({"no_limiter_" + k: v for k, v in d_plain.items()})
    np.savez_compressed(os.path.join(OUT, "c3_pipeline_96k.npz"), input_sum=input_sum, target_seed=0, reference_seed=1,
                        sample_rate=96000, **out)


def main():
    warnings.simplefilter("ignore")
    m = import_reference()
    os.makedirs(OUT, exist_ok=True)
    which = [a for a in sys.argv[1:] if a in ("c3", "c5")] or ["c5", "c3"]
    for w in which:
        {"c3": make_c3, "c5": make_c5}[w](m)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
