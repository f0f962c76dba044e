"""Numpy models of what the CUDA kernels compute, step for step (tests only).

They exist to check the ALGORITHM the kernels implement (plan tables, block-parallel
substitution sweeps with a warm-up, overlap-save framing) against the oracle before any GPU
time is spent, and to localise a failing GPU parity test to one stage.
"""
import numpy as np

WARM = 48  # rows of warm-up before a thread's block of the substitution sweeps


def spline_moments(y, hinv, lu, end, block=16, warm=WARM):
    """Second derivatives at all knots: rhs, forward + backward sweep done block-parallel."""
    n = len(y)
    m = n - 2
    delta = (y[1:] - y[:-1]) * hinv
    d = 6.0 * (delta[1:] - delta[:-1])
    fa, invden, cp = lu
    z = np.zeros(m)
    for start in range(0, m, block):
        lo = max(0, start - warm)
        acc = 0.0
        for i in range(lo, min(m, start + block)):
            acc = d[i] * invden[i] - fa[i] * acc
            if i >= start:
                z[i] = acc
    M = np.zeros(m)
    for start in range(0, m, block):
        hi = min(m - 1, start + block - 1 + warm)
        acc = 0.0
        for i in range(hi, start - 1, -1):
            acc = z[i] - cp[i] * acc
            if i < start + block:
                M[i] = acc
    full = np.empty(n)
    full[1:-1] = M
    full[0] = end[0] * M[0] + end[1] * M[1]
    full[-1] = end[2] * M[-1] + end[3] * M[-2]
    return full


def spline_eval(y, M, idx, w):
    return w[:, 0] * y[idx] + w[:, 1] * y[idx + 1] + w[:, 2] * M[idx] + w[:, 3] * M[idx + 1]


def lowess_fit(y, a, k):
    fits = np.array([a["lw_rows"][r] @ y[l:l + k] for r, l in zip(a["lw_row_idx"], a["lw_fit_left"])])
    seg, al = a["lw_seg"], a["lw_alpha"]
    nxt = np.minimum(seg + 1, len(fits) - 1)
    return al * fits[nxt] + (1.0 - al) * fits[seg]


def design_fir(avg_t, avg_r, tables):
    """FIR of one channel from scaled average spectra, the way design.cu does it."""
    a = tables.arrays
    m = avg_r / np.maximum(tables.min_value, avg_t)
    M1 = spline_moments(m, a["sa_hinv"], a["sa_lu"], a["sa_end"])
    m_log = spline_eval(m, M1, a["sa_eval_idx"], a["sa_eval_w"])
    s_log = lowess_fit(m_log, a, tables.lowess_k)
    M2 = spline_moments(s_log, a["sb_hinv"], a["sb_lu"], a["sb_end"])
    s = spline_eval(s_log, M2, a["sb_eval_idx"], a["sb_eval_w"])
    s[0] = 0.0
    s[1] = m[1]
    F = tables.fft_size
    spec = np.concatenate([s, s[-2:0:-1]])
    h = np.fft.ifft(spec).real
    fir = np.roll(h, F // 2) * a["hann"]
    return fir
