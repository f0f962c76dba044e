"""oracle/lowess.py against the mathematical definition (statsmodels itself is not installable
here: parity unpinned for this one function, see its header)."""
import numpy as np

import lowess as lw


def brute_force(y, x, frac):
    n = len(x)
    k = int(frac * n + 1e-10)
    out = np.empty(n)
    for i in range(n):
        order = np.argsort(np.abs(x - x[i]), kind="stable")[:k]
        lo, hi = order.min(), order.max()
        # nearest-neighbour window as statsmodels slides it (ties resolved to the left window)
        xs, ys = x[lo:hi + 1], y[lo:hi + 1]
        d = np.abs(xs - x[i])
        radius = max(d[0], d[-1])
        w = (1 - (d / radius) ** 3) ** 3
        w[d >= radius] = 0
        W = np.diag(w)
        A = np.stack([np.ones_like(xs), xs], axis=1)
        beta = np.linalg.solve(A.T @ W @ A, A.T @ W @ ys)
        out[i] = beta[0] + beta[1] * x[i]
    return out


def test_reproduces_straight_lines():
    x = np.linspace(0, 1, 501)
    y = 3.0 - 2.5 * x
    for delta in (0.0, 0.01):
        assert np.abs(lw.lowess(y, x, 0.1, 0, delta) - y).max() < 1e-12


def test_matches_weighted_least_squares_at_delta_zero():
    rng = np.random.default_rng(0)
    x = np.linspace(0, 1, 301)
    y = np.sin(6 * x) + 0.1 * rng.standard_normal(301)
    got = lw.lowess(y, x, 0.08, 0, 0.0)
    want = brute_force(y, x, 0.08)
    # interior windows are unambiguous; edges depend on the tie rule of the sliding window
    assert np.abs(got - want)[30:-30].max() < 1e-11


def test_delta_skips_and_interpolates():
    rng = np.random.default_rng(1)
    n = 8193
    x = np.linspace(0, 1, n)
    y = np.cumsum(rng.standard_normal(n)) / 50
    fit_idx, left, k = lw.lowess_plan(x, 0.0375, 0.001)
    assert k == 307
    assert list(fit_idx[:3]) == [0, 8, 16] and list(fit_idx[-3:]) == [8184, 8191, 8192]
    full = lw.lowess(y, x, 0.0375, 0, 0.0)
    skipped = lw.lowess(y, x, 0.0375, 0, 0.001)
    assert np.abs(full[fit_idx] - skipped[fit_idx]).max() < 1e-13
    j = 12  # between regression points 8 and 16
    a = (x[j] - x[8]) / (x[16] - x[8])
    assert abs(skipped[j] - (a * skipped[16] + (1 - a) * skipped[8])) < 1e-15


def test_robust_iterations_run():
    rng = np.random.default_rng(2)
    x = np.linspace(0, 1, 200)
    y = x + 0.01 * rng.standard_normal(200)
    y[50] += 5.0
    plain = lw.lowess(y, x, 0.2, 0, 0.0)
    robust = lw.lowess(y, x, 0.2, 2, 0.0)
    assert abs(robust[50] - x[50]) < abs(plain[50] - x[50])
