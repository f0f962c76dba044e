"""ORACLE (test infrastructure only) -- LOWESS as statsmodels computes it.

The reference calls ``statsmodels.api.nonparametric.lowess`` at
``matchering/dsp.py:103-106``.  statsmodels is a third-party dependency that is
NOT vendored under /root/reference and NOT installed in this image
(``requirements.txt:5`` pins only ``statsmodels>=0.13.2``), so this file restates
the published algorithm of ``statsmodels/nonparametric/_smoothers_lowess.pyx``
(Cleveland 1979 with the `delta` skipping of the netlib `lowess.f`).

PARITY UNPINNED for this one function: there is no statsmodels binary here to
check it against and the reference ships no golden vectors.  It is validated
against the mathematical definition only (tests/test_oracle_lowess.py: exact
reproduction of straight lines, agreement with a brute-force tricube-weighted
least-squares fit at delta=0, and linear interpolation between the delta-skipped
fits).

Nothing in the product path (``matchering_b200``) may import this file.
"""
import numpy as np


def lowess_plan(x: np.ndarray, frac: float, delta: float):
    """Index bookkeeping of one LOWESS pass; depends on the abscissa only.

    Returns (fit_idx, left, k): the indices at which a local regression is
    evaluated, the left edge of each regression's neighbourhood [left, left+k),
    following update_neighborhood / update_indices of the statsmodels kernel.
    """
    n = len(x)
    k = int(frac * n + 1e-10)
    if not 2 <= k <= n:
        raise ValueError("lowess: frac must give 2 <= k <= n")
    fit_idx, lefts = [], []
    i, last_fit, left, right = 0, -1, 0, k
    while True:
        # update_neighborhood: slide right while the left edge is the farther one
        while right < n and (x[i] - x[left]) > (x[right] - x[i]):
            left += 1
            right += 1
        fit_idx.append(i)
        lefts.append(left)
        # update_indices
        last_fit = i
        cut = x[last_fit] + delta
        kk = last_fit
        for kk in range(last_fit + 1, n):
            if x[kk] > cut:
                break
            if x[kk] == x[last_fit]:
                # tie: statsmodels copies the fit; never happens on a linspace grid
                fit_idx.append(kk)
                lefts.append(left)
                last_fit = kk
        i = max(kk - 1, last_fit + 1)
        if last_fit >= n - 1:
            break
    return np.asarray(fit_idx, dtype=np.int64), np.asarray(lefts, dtype=np.int64), k


def lowess(y: np.ndarray, x: np.ndarray, frac: float, it: int, delta: float) -> np.ndarray:
    """Fitted values only (column 1 of statsmodels' return), x sorted ascending."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    fit_idx, lefts, k = lowess_plan(x, frac, delta)
    resid_w = np.ones(n)
    y_fit = np.zeros(n)
    for robust_iter in range(it + 1):
        last = -1
        for i, left in zip(fit_idx, lefts):
            if last >= 0 and x[i] == x[last]:
                y_fit[i] = y_fit[last]
                last = i
                continue
            xs = x[left:left + k]
            dist = np.abs(xs - x[i])
            radius = max(dist[0], dist[-1])
            w = dist / radius
            w = (1.0 - w * w * w)
            w = w * w * w
            w[dist >= radius] = 0.0
            w = w * resid_w[left:left + k]
            sw = w.sum()
            if sw <= 0.0 or np.count_nonzero(w) == 1:
                y_fit[i] = y[i]
            else:
                w = w / sw
                xbar = np.sum(w * xs)
                sqdev = np.sum(w * (xs - xbar) ** 2)
                p = w * (1.0 + (x[i] - xbar) * (xs - xbar) / sqdev)
                y_fit[i] = np.sum(p * y[left:left + k])
            if last < i - 1 and last >= 0:
                denom = x[i] - x[last]
                a = (x[last + 1:i] - x[last]) / denom
                y_fit[last + 1:i] = a * y_fit[i] + (1.0 - a) * y_fit[last]
            last = i
        if robust_iter < it:
            # calculate_residual_weights: |residual| / (6 * median), trimmed at 1, through the bisquare;
            # a zero median keeps weight 1 on the exact fits and 0 elsewhere
            u = np.abs(y - y_fit)
            s = np.median(u)
            u = np.where(u > 0, 1.0, 0.0) if s == 0 else u / (6.0 * s)
            u = np.minimum(u, 1.0)
            resid_w = (1.0 - u * u) ** 2
    return y_fit
