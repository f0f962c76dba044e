"""ORACLE (test infrastructure only) -- CPU restatement of Matchering's DSP hot path.

A float64 numpy/scipy restatement of sergree/matchering v2.0.6 ``stages.main`` and
``limiter.limit``; every function cites the reference file:line it follows.  It is
written from the *semantics* (SURVEY.md Appendix A), not transliterated: the STFT is
an explicit frame reshape + rfft, ``filtfilt`` is two explicit one-pole passes over an
explicit odd extension, the sliding windows are explicit index ranges.  That makes
every identity the CUDA kernels rely on a tested statement (tests/test_oracle_*.py
pin this file against the unmodified reference run in the build container, and
against the golden vectors committed under tests/golden/).

Pinning status: pinned against the real reference (imported through
oracle/ref_shims.py) for everything except LOWESS, whose statsmodels original is not
available in this image -- see oracle/lowess.py ("parity unpinned" for that one call).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (matchering_b200) never does.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
from scipy import fft as _fft
from scipy import interpolate as _interp
from scipy import signal as _signal

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import lowess as _lowess  # noqa: E402  (oracle/lowess.py)


# --------------------------------------------------------------------------- config
class OracleLimiterConfig:
    """Defaults of matchering/defaults.py:25-58."""

    def __init__(self, attack=1.0, hold=1.0, release=3000.0, attack_filter_coefficient=-2.0,
                 hold_filter_order=1, hold_filter_coefficient=7.0,
                 release_filter_order=1, release_filter_coefficient=800.0):
        self.attack = attack
        self.hold = hold
        self.release = release
        self.attack_filter_coefficient = attack_filter_coefficient
        self.hold_filter_order = hold_filter_order
        self.hold_filter_coefficient = hold_filter_coefficient
        self.release_filter_order = release_filter_order
        self.release_filter_coefficient = release_filter_coefficient


class OracleConfig:
    """The hot-path subset of matchering/defaults.py:61-155 (max_piece_size in samples)."""

    def __init__(self, internal_sample_rate=44100, max_piece_size=15.0,
                 threshold=(2 ** 15 - 61) / 2 ** 15, min_value=1e-6, fft_size=4096,
                 lin_log_oversampling=4, rms_correction_steps=4, lowess_frac=0.0375,
                 lowess_it=0, lowess_delta=0.001, limiter=None):
        self.internal_sample_rate = internal_sample_rate
        self.max_piece_size = max_piece_size * internal_sample_rate  # defaults.py:109
        self.threshold = threshold
        self.min_value = min_value
        self.fft_size = fft_size
        self.lin_log_oversampling = lin_log_oversampling
        self.rms_correction_steps = rms_correction_steps
        self.lowess_frac = lowess_frac
        self.lowess_it = lowess_it
        self.lowess_delta = lowess_delta
        self.limiter = limiter if limiter is not None else OracleLimiterConfig()


def config_from(cfg) -> "OracleConfig":
    """Accept an OracleConfig, a reference Config or a matchering_b200 Config."""
    if isinstance(cfg, OracleConfig):
        return cfg
    out = OracleConfig.__new__(OracleConfig)
    for name in ("internal_sample_rate", "max_piece_size", "threshold", "min_value", "fft_size",
                 "lin_log_oversampling", "rms_correction_steps", "lowess_frac", "lowess_it",
                 "lowess_delta"):
        setattr(out, name, getattr(cfg, name))
    lim = OracleLimiterConfig()
    for name in vars(lim):
        setattr(lim, name, getattr(cfg.limiter, name))
    out.limiter = lim
    return out


# --------------------------------------------------------------------------- levels
def mid_side(x: np.ndarray):
    """dsp.py:57-64: mid = (L+R)*0.5, side = mid - R."""
    mid = (x[:, 0] + x[:, 1]) * 0.5
    return mid, mid - x[:, 1]


def piece_layout(n: int, max_piece_size: float):
    """stage_helpers/match_levels.py:47-59 (float division then int truncation)."""
    divisions = int(n / max_piece_size) + 1
    piece = int(n / divisions)
    return divisions, piece


def piece_rms(v: np.ndarray, piece: int, divisions: int) -> np.ndarray:
    """dsp.py:71-86: sqrt(sum(x^2)/piece) over the first piece*divisions samples."""
    u = v[: piece * divisions].reshape(divisions, piece)
    return np.sqrt(np.einsum("ij,ij->i", u, u) / piece)


def loudest_mask_and_match_rms(rmses: np.ndarray):
    """match_levels.py:62-71,93-103: avg = rms(rmses); mask = rmses >= avg; rms of the masked."""
    avg = math.sqrt(float(rmses @ rmses) / rmses.shape[0])
    mask = rmses >= avg
    sel = rmses[mask]
    match = math.sqrt(float(sel @ sel) / sel.shape[0])
    return mask, match, avg


def normalize(x: np.ndarray, threshold: float, eps: float, normalize_clipped: bool):
    """dsp.py:93-100."""
    peak = float(np.abs(x).max())
    coef = 1.0
    if peak < threshold or normalize_clipped:
        coef = max(eps, peak / threshold)
    return x / coef, coef


# --------------------------------------------------------------------------- spectra / FIR
def average_spectrum(v: np.ndarray, piece: int, mask: np.ndarray, fft_size: int) -> np.ndarray:
    """match_frequencies.py:30-42 with scipy stft(boxcar, noverlap=0, boundary=None,
    padded=False): per selected piece, floor(piece/F) contiguous frames, rfft/F, |.|, mean."""
    divisions = mask.shape[0]
    frames = piece // fft_size
    u = v[: piece * divisions].reshape(divisions, piece)[mask][:, : frames * fft_size]
    u = u.reshape(-1, frames, fft_size)
    spec = np.abs(_fft.rfft(u, axis=-1)) / fft_size
    return spec.mean(axis=(0, 1))


def frequency_grids(cfg: OracleConfig):
    """match_frequencies.py:46-58."""
    half = cfg.fft_size // 2
    lin = cfg.internal_sample_rate * 0.5 * np.linspace(0, 1, half + 1)
    log = cfg.internal_sample_rate * 0.5 * np.logspace(
        np.log10(4 / cfg.fft_size), 0, half * cfg.lin_log_oversampling + 1)
    return lin, log


def smooth_matching_curve(m: np.ndarray, cfg: OracleConfig) -> np.ndarray:
    """match_frequencies.py:45-75: not-a-knot cubic lin->log, LOWESS on an index-uniform
    abscissa (dsp.py:103-106), not-a-knot cubic log->lin with extrapolation, then the
    two overrides."""
    lin, log = frequency_grids(cfg)
    m_log = _interp.make_interp_spline(lin, m, k=3)(log)
    s_log = _lowess.lowess(m_log, np.linspace(0, 1, len(m_log)), cfg.lowess_frac,
                           cfg.lowess_it, cfg.lowess_delta)
    s = _interp.make_interp_spline(log, s_log, k=3)(lin)  # BSpline extrapolates by default
    s[0] = 0.0
    s[1] = m[1]
    return s


def design_fir(avg_target: np.ndarray, avg_reference: np.ndarray, cfg: OracleConfig) -> np.ndarray:
    """match_frequencies.py:78-101."""
    m = avg_reference / np.maximum(cfg.min_value, avg_target)
    s = smooth_matching_curve(m, cfg)
    h = np.fft.irfft(s)
    return np.fft.ifftshift(h) * _signal.windows.hann(len(h))


def convolve_same(x: np.ndarray, fir: np.ndarray) -> np.ndarray:
    """match_frequencies.py:112-113: fftconvolve(x, fir, 'same') = full[(F-1)//2 : (F-1)//2+N]."""
    full = _signal.fftconvolve(x, fir, "full")
    start = (len(fir) - 1) // 2
    return full[start:start + len(x)]


# --------------------------------------------------------------------------- limiter
def rectified_gain(x: np.ndarray, threshold: float):
    """dsp.py:117-121 and hyrax.py:87: r = max(max|L|,|R|, thr)/thr ; g = 1 - 1/r."""
    r = np.maximum(np.abs(x).max(axis=1), threshold) / threshold
    return r, 1.0 - 1.0 / r


def centred_max(g: np.ndarray, reach: int) -> np.ndarray:
    """hyrax.py:35-37: maximum_filter1d(size=2w-1, mode='reflect') == max over the valid
    part of [n-reach, n+reach] (the reflected samples are already inside the window)."""
    n = len(g)
    pad = np.full(reach, -np.inf)
    p = np.concatenate([pad, g, pad])
    win = np.lib.stride_tricks.sliding_window_view(p, 2 * reach + 1)
    return win.max(axis=1)[:n] if n < 200000 else _blocked_window_max(p, 2 * reach + 1, n)


def trailing_max(a: np.ndarray, length: int) -> np.ndarray:
    """hyrax.py:38-40: zero left-pad + maximum_filter1d + trim == max(a[n-length+1 .. n]),
    indices below zero contributing 0."""
    n = len(a)
    p = np.concatenate([np.zeros(length - 1), a])
    win = np.lib.stride_tricks.sliding_window_view(p, length)
    return win.max(axis=1)[:n] if n < 200000 else _blocked_window_max(p, length, n)


def _blocked_window_max(p: np.ndarray, length: int, n: int) -> np.ndarray:
    out = np.empty(n)
    step = 1 << 16
    for s in range(0, n, step):
        e = min(n, s + step)
        win = np.lib.stride_tricks.sliding_window_view(p[s:e + length - 1], length)
        out[s:e] = win.max(axis=1)
    return out


def one_pole_forward_backward(a_in: np.ndarray, c: float) -> np.ndarray:
    """hyrax.py:48-51 filtfilt(b=[1-c], a=[1,-c]) per scipy/signal/_signaltools.py filtfilt:
    odd extension by 6, steady-state initial condition zi = c scaled by the first sample,
    forward pass, then the same on the reversed signal, trim the extension."""
    edge = 6
    if len(a_in) <= edge:
        raise ValueError("The length of the input vector x must be greater than padlen, which is 6.")
    ext = np.concatenate([2 * a_in[0] - a_in[edge:0:-1], a_in, 2 * a_in[-1] - a_in[-2:-edge - 2:-1]])
    b, a = [1.0 - c], [1.0, -c]
    f, _ = _signal.lfilter(b, a, ext, zi=[c * ext[0]])
    r, _ = _signal.lfilter(b, a, f[::-1], zi=[c * f[-1]])
    return r[::-1][edge:-edge]


def limiter_coefficients(cfg: OracleConfig):
    """utils.py:50-55, hyrax.py:44-48,57-72."""
    sr = cfg.internal_sample_rate
    lim = cfg.limiter
    attack = int(sr * lim.attack * 1e-3)
    hold = int(sr * lim.hold * 1e-3)
    reach = (attack + 1 if not attack & 1 else attack) - 1  # make_odd(attack) - 1
    c = math.exp(lim.attack_filter_coefficient / attack)
    bh, ah = _signal.butter(lim.hold_filter_order, lim.hold_filter_coefficient, fs=sr)
    br, ar = _signal.butter(lim.release_filter_order, lim.release_filter_coefficient / lim.release, fs=sr)
    return dict(attack=attack, hold=hold, reach=reach, c=c, bh=bh, ah=ah, br=br, ar=ar)


def limit(x: np.ndarray, cfg, trace: dict | None = None) -> np.ndarray:
    """limiter/hyrax.py:78-99."""
    cfg = config_from(cfg)
    r, g = rectified_gain(x, cfg.threshold)
    if np.all(np.isclose(r, 1.0)):
        return x
    k = limiter_coefficients(cfg)
    a_env = centred_max(g, k["reach"])
    g_att = one_pole_forward_backward(a_env, k["c"])
    h_env = trailing_max(a_env, k["hold"])
    hold_out = _signal.lfilter(k["bh"], k["ah"], h_env)
    rel_out = _signal.lfilter(k["br"], k["ar"], np.maximum(h_env, hold_out))
    g_rel = np.maximum(hold_out, rel_out)
    gain = 1.0 - np.maximum(np.maximum(g, g_att), g_rel)
    if trace is not None:
        trace.update(g=g, a_env=a_env, g_att=g_att, h_env=h_env, hold_out=hold_out,
                     rel_out=rel_out, gain=gain)
    return x * gain[:, None]


# --------------------------------------------------------------------------- pipeline
def analyze(x: np.ndarray, cfg: OracleConfig):
    """match_levels.py:134-161 without the piece gather (a mask replaces it)."""
    mid, side = mid_side(x)
    divisions, piece = piece_layout(len(mid), cfg.max_piece_size)
    rmses = piece_rms(mid, piece, divisions)
    mask, match_rms, avg = loudest_mask_and_match_rms(rmses)
    return dict(mid=mid, side=side, divisions=divisions, piece=piece, rmses=rmses,
                mask=mask, match_rms=match_rms, avg=avg)


def main(target: np.ndarray, reference: np.ndarray, cfg, need_default=True,
         need_no_limiter=False, need_no_limiter_normalized=False, trace: dict | None = None):
    """stages.py:210-272 (same returns: result, result_no_limiter, result_no_limiter_normalized)."""
    cfg = config_from(cfg)
    F = cfg.fft_size
    # __match_levels, stages.py:38-104
    reference, final_coef = normalize(reference, cfg.threshold, cfg.min_value, False)
    ta = analyze(target, cfg)
    ra = analyze(reference, cfg)
    c0 = ra["match_rms"] / max(cfg.min_value, ta["match_rms"])
    t_mid, t_side = ta["mid"] * c0, ta["side"] * c0
    # __match_frequencies, stages.py:107-135 ; the loudest target pieces are scaled by c0 too
    firs, avgs = {}, {}
    for name, tv, rv in (("mid", t_mid, ra["mid"]), ("side", t_side, ra["side"])):
        at = average_spectrum(tv, ta["piece"], ta["mask"], F)
        ar = average_spectrum(rv, ra["piece"], ra["mask"], F)
        avgs[name] = (at, ar)
        firs[name] = design_fir(at, ar, cfg)
    r_mid = convolve_same(t_mid, firs["mid"])
    r_side = convolve_same(t_side, firs["side"])
    result = np.stack([r_mid + r_side, r_mid - r_side], axis=1)  # dsp.py:67-68
    # __correct_levels, stages.py:138-170
    coefs = []
    for _ in range(cfg.rms_correction_steps):
        rm = piece_rms(np.clip(r_mid, -1.0, 1.0), ta["piece"], ta["divisions"])
        _, match, _ = loudest_mask_and_match_rms(rm)
        c = ra["match_rms"] / max(cfg.min_value, match)
        coefs.append(c)
        r_mid = r_mid * c
        result = result * c
    if trace is not None:
        trace.update(final_coef=final_coef, target=ta, reference=ra, c0=c0, firs=firs, avgs=avgs,
                     correction=coefs, pre_limiter=result)
    # __finalize, stages.py:173-207
    out_norm = None
    if need_no_limiter_normalized:
        out_norm, _ = normalize(result, cfg.threshold, cfg.min_value, True)
    out = None
    if need_default:
        out = limit(result, cfg) * final_coef
    return out, (result if need_no_limiter else None), out_norm


# --------------------------------------------------------------------------- preview creator
def preview_pieces(target: np.ndarray, result: np.ndarray, cfg):
    """preview_creator.create_preview (matchering/preview_creator.py:30-94) up to the two arrays it
    saves: windows of preview_size every preview_analysis_step (dsp.strided_app_2d, dsp.py:128-141),
    the window where the RESULT is loudest (dsp.batch_rms_2d, dsp.py:144-145), the same window of
    the target clipped at the threshold, both with linear fades (dsp.fade, dsp.py:148-155) unless
    the window is the whole track.  -> (window index, target piece, result piece)"""
    size, step = int(cfg.preview_size), int(cfg.preview_analysis_step)
    target = np.clip(np.asarray(target, dtype=np.float64), -cfg.threshold, cfg.threshold)
    result = np.asarray(result, dtype=np.float64)
    n = result.shape[0]
    if size > n:
        index, lo, hi = 0, 0, n
    else:
        count = (n - size) // step + 1
        energy = np.array([np.sum(result[w * step:w * step + size] ** 2) for w in range(count)])
        index = int(np.argmax(np.sqrt(energy / (2 * size))))
        lo, hi = index * step, index * step + size
    t_piece, r_piece = target[lo:hi].copy(), result[lo:hi].copy()
    if hi - lo != n:
        fade = int(min(cfg.preview_fade_size, (hi - lo) // cfg.preview_fade_coefficient))
        ramp = np.linspace(0, 1, fade)
        for piece in (t_piece, r_piece):
            piece[:fade] *= ramp[:, None]
            piece[len(piece) - fade:] *= ramp[::-1, None]
    return index, t_piece, r_piece


# --------------------------------------------------------------------------- synthetic inputs
_PINK_B = [0.049922035, -0.095993537, 0.050612699, -0.004408786]
_PINK_A = [1.0, -2.494956002, 2.017265875, -0.522189400]


def synth_target(n: int, seed: int, kind: str = "enveloped") -> np.ndarray:
    """white noise (optionally under a slow |sin| envelope);
    generated in float32 so the GPU and the float64 oracle see identical sample values."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.5, 0.5, (n, 2))
    if kind == "enveloped":
        x = x * (0.25 + 0.75 * np.abs(np.sin(np.linspace(0, 40, n))))[:, None]
    return np.ascontiguousarray(x.astype(np.float32))


def synth_reference(n: int, seed: int, kind: str = "loud") -> np.ndarray:
    """pink noise, either peak-normalised to 0.9 or tanh-compressed."""
    rng = np.random.default_rng(seed)
    x = _signal.lfilter(_PINK_B, _PINK_A, rng.standard_normal((n, 2)), axis=0)
    x = x / np.abs(x).max()
    x = np.tanh(3.0 * x) if kind == "loud" else 0.9 * x
    return np.ascontiguousarray(x.astype(np.float32))


def synth_limiter_input(n: int, seed: int) -> np.ndarray:
    """SURVEY.md section 8(d) C5 recipe."""
    rng = np.random.default_rng(seed)
    x = (rng.uniform(0, 1, (n, 2)) * 3 - 1.5) * (0.3 + 0.7 * np.abs(np.sin(np.linspace(0, 500, n))))[:, None]
    return np.ascontiguousarray(x.astype(np.float32))


def synth_limiter_input_prefix(n: int, m: int, seed: int) -> np.ndarray:
    """The first m frames of synth_limiter_input(n, seed) without synthesising the other n - m (the
    generator fills row-major, and linspace(0, 500, n)[k] = k * 500 / (n - 1))."""
    rng = np.random.default_rng(seed)
    env = 0.3 + 0.7 * np.abs(np.sin(np.linspace(0, 500, n)[:m] if n < (1 << 22) else np.arange(m) * (500.0 / (n - 1))))
    x = (rng.uniform(0, 1, (m, 2)) * 3 - 1.5) * env[:, None]
    return np.ascontiguousarray(x.astype(np.float32))


def synth_tonal(n: int, seed: int, sample_rate: int = 44100) -> np.ndarray:
    """Decaying harmonic notes with random panning between stretches of exact digital silence
    (head and tail): pieces and frames with zero energy, strongly coloured spectra."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sample_rate
    x = np.zeros((n, 2))
    for _ in range(12):
        f0, start, dur, pan = rng.uniform(60, 2000), rng.uniform(0.1, t[-1] - 0.2), rng.uniform(0.05, 0.6), rng.uniform(0, 1)
        env = np.where(t >= start, np.exp(-(t - start) / dur), 0.0)
        tone = sum(np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6)) / h for h in range(1, 6))
        x[:, 0] += pan * env * tone * 0.2
        x[:, 1] += (1 - pan) * env * tone * 0.2
    x[: int(0.1 * sample_rate)] = 0.0
    x[n - int(0.05 * sample_rate):] = 0.0
    return np.ascontiguousarray(x.astype(np.float32))
