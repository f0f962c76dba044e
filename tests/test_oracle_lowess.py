"""oracle/lowess.py against the mathematical definition (statsmodels itself is not installable
here: parity unpinned for this one function, see its header)."""
import numpy as np

import lowess as lw


def brute_force(y, x, frac, resid_w=None):
    n = len(x)
    k = int(frac * n + 1e-10)
    out = np.empty(n)
    rw = np.ones(n) if resid_w is None else resid_w
    for i in range(n):
        order = np.argsort(np.abs(x - x[i]), kind="stable")[:k]
        lo, hi = order.min(), order.max()
        # nearest-neighbour window as statsmodels slides it (ties resolved to the left window)
        xs, ys = x[lo:hi + 1], y[lo:hi + 1]
        d = np.abs(xs - x[i])
        radius = max(d[0], d[-1])
        w = (1 - (d / radius) ** 3) ** 3
        w[d >= radius] = 0
        W = np.diag(w * rw[lo:hi + 1])
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


def test_robust_iterations_match_reweighted_least_squares():
    """Two robustness iterations against brute-force weighted least squares re-weighted by the bisquare of
    the residuals over 6 * median (Cleveland 1979; statsmodels calculate_residual_weights)."""
    rng = np.random.default_rng(3)
    x = np.linspace(0, 1, 301)
    y = np.sin(6 * x) + 0.1 * rng.standard_normal(301)
    y[[40, 150, 220]] += np.array([3.0, -4.0, 2.5])
    got = lw.lowess(y, x, 0.08, 2, 0.0)
    rw, fit = None, None
    for _ in range(3):
        fit = brute_force(y, x, 0.08, rw)
        # the edges of the brute-force fit follow a different tie rule; take the oracle's values there so
        # that the residual weights of the interior are built from the same numbers
        fit[:30], fit[-30:] = lw.lowess(y, x, 0.08, _, 0.0)[:30], lw.lowess(y, x, 0.08, _, 0.0)[-30:]
        u = np.abs(y - fit)
        u = np.minimum(u / (6.0 * np.median(u)), 1.0)
        rw = (1 - u * u) ** 2
    assert np.abs(got - fit)[60:-60].max() < 1e-10
