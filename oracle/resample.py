"""ORACLE (test infrastructure only) -- resampy.resample(x, sr_orig, sr_new, axis=0) as the reference calls
it at matchering/checker.py:30-44 (filter "kaiser_best", the default).

resampy is a third-party dependency (requirements.txt:4, ``resampy>=0.4.2``) that is neither vendored under
/root/reference nor installable in this image, and the reference ships no test that pins its output:
PARITY UNPINNED for this function.  This file restates resampy 0.4's published algorithm
(``resampy/core.py`` resample, ``resampy/interpn.py`` _resample_loop, ``resampy/filters.py`` sinc_window):

* the interpolation table is half of a Kaiser-windowed sinc, ``num_zeros = 64`` zero crossings sampled
  ``2**9 = 512`` times each, ``rolloff = 0.9475937167399596``, Kaiser ``beta = 14.769656459379492``
  (the parameters documented for ``kaiser_best``; the package ships the table precomputed from them);
* output sample t sits at input time ``t / ratio``; its value is the sum over the input samples left and right
  of that time of ``x * (win[offset + i*step] + eta * delta[offset + i*step])`` -- the table read at ``step =
  int(scale*512)`` entries per input sample with linear interpolation between entries; when downsampling the
  table is scaled by the ratio and stretched (``scale = ratio``) so that it cuts off below the new Nyquist;
* ``len(y) = int(len(x) * sr_new / sr_orig)``.

It is validated against properties only (tests/test_oracle_resample.py: identity at equal rates up to the
roll-off, a sine keeps its frequency and amplitude in the pass band, content above the new Nyquist is removed).
Nothing in the product path may import this file.
"""
import numpy as np
from scipy.signal.windows import kaiser

NUM_ZEROS = 64
PRECISION = 9
ROLLOFF = 0.9475937167399596
BETA = 14.769656459379492


def kaiser_best_filter():
    """resampy.filters.sinc_window(num_zeros=64, precision=9, window=kaiser(beta), rolloff) ->
    (half window float64 [32769], samples per zero crossing)."""
    num_bits = 2 ** PRECISION
    n = num_bits * NUM_ZEROS
    sinc_win = ROLLOFF * np.sinc(ROLLOFF * np.linspace(0, NUM_ZEROS, num=n + 1, endpoint=True))
    taper = kaiser(2 * n + 1, BETA)[n:]
    return taper * sinc_win, num_bits


def resample(x: np.ndarray, sr_orig: int, sr_new: int) -> np.ndarray:
    """(frames, channels) float64 -> (int(frames * sr_new / sr_orig), channels) float64."""
    x = np.asarray(x, dtype=np.float64)
    sample_ratio = float(sr_new) / sr_orig
    n_orig = x.shape[0]
    n_out = int(n_orig * sr_new / sr_orig)
    interp_win, num_table = kaiser_best_filter()
    if sample_ratio < 1:
        interp_win = sample_ratio * interp_win
    interp_delta = np.diff(interp_win, append=interp_win[-1])
    scale = min(1.0, sample_ratio)
    t_out = np.arange(n_out) * (1.0 / sample_ratio)
    index_step = int(scale * num_table)
    nwin = interp_win.shape[0]
    y = np.zeros((n_out,) + x.shape[1:], dtype=np.float64)

    n = t_out.astype(np.int64)
    frac = scale * (t_out - n)
    # left wing: x[n - i], i = 0 .. i_max-1
    index_frac = frac * num_table
    offset = index_frac.astype(np.int64)
    eta = index_frac - offset
    i_max = np.minimum(n + 1, (nwin - offset) // index_step)
    for i in range(int(i_max.max()) if n_out else 0):
        live = i < i_max
        idx = np.where(live, offset + i * index_step, 0)
        weight = np.where(live, interp_win[idx] + eta * interp_delta[idx], 0.0)
        y += weight[:, None] * x[np.where(live, n - i, 0)]
    # right wing: x[n + k + 1], k = 0 .. k_max-1
    frac = scale - frac
    index_frac = frac * num_table
    offset = index_frac.astype(np.int64)
    eta = index_frac - offset
    k_max = np.minimum(n_orig - n - 1, (nwin - offset) // index_step)
    for k in range(int(k_max.max()) if n_out else 0):
        live = k < k_max
        idx = np.where(live, offset + k * index_step, 0)
        weight = np.where(live, interp_win[idx] + eta * interp_delta[idx], 0.0)
        y += weight[:, None] * x[np.where(live, n + k + 1, 0)]
    return y
