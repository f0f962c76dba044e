"""Resampling to Config.internal_sample_rate on the device (reference: matchering/checker.py:30-44, which
calls resampy.resample(array, sr, required_sr, axis=0) with its default filter "kaiser_best").

The interpolation table is Config-free: half of a Kaiser-windowed sinc with resampy's documented kaiser_best
parameters (64 zero crossings, 2**9 table entries per crossing, roll-off 0.9475937167399596, Kaiser beta
14.769656459379492; resampy/filters.py sinc_window).  It is built here once, in float64, scaled by the ratio
when downsampling (resampy/core.py), paired with its forward differences and uploaded; the kernel
(csrc/resample.cu, mgb_resample) does the per-sample work.  resampy itself is not installable in this image,
so its exact table cannot be compared: parity unpinned for this step.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from scipy.signal.windows import kaiser

from . import _native

NUM_ZEROS, PRECISION = 64, 9
ROLLOFF, BETA = 0.9475937167399596, 14.769656459379492

_TABLES: dict = {}


def _table(sample_ratio: float, device):
    key = (device.index, sample_ratio if sample_ratio < 1 else 1.0)
    got = _TABLES.get(key)
    if got is None:
        num_bits = 2 ** PRECISION
        n = num_bits * NUM_ZEROS
        win = ROLLOFF * np.sinc(ROLLOFF * np.linspace(0, NUM_ZEROS, num=n + 1, endpoint=True)) * kaiser(2 * n + 1, BETA)[n:]
        if sample_ratio < 1:
            win = sample_ratio * win
        delta = np.diff(win, append=win[-1])
        pairs = np.ascontiguousarray(np.stack([win, delta], axis=1))  # (window, forward difference) per entry: one 16-byte load
        got = _TABLES[key] = (torch.from_numpy(pairs).to(device), len(win), num_bits)
    return got


def resample_on_device(audio: torch.Tensor, rate_in: int, rate_out: int) -> torch.Tensor:
    """(frames, 2) float32 CUDA tensor at rate_in -> (int(frames * rate_out / rate_in), 2) at rate_out."""
    lib = _native.load()
    assert audio.is_cuda and audio.dtype == torch.float32 and audio.ndim == 2 and audio.shape[1] == 2
    audio = audio.contiguous()
    frames_in = audio.shape[0]
    frames_out = int(lib.mgb_resample_frames(frames_in, rate_in, rate_out))
    table, nwin, num_table = _table(float(rate_out) / rate_in, audio.device)
    out = torch.empty((frames_out, 2), dtype=torch.float32, device=audio.device)
    stream = C.c_void_p(torch.cuda.current_stream(audio.device).cuda_stream)
    with torch.cuda.device(audio.device):
        _native.check(lib, lib.mgb_resample(audio.data_ptr(), frames_in, rate_in, out.data_ptr(), frames_out, rate_out,
                                            table.data_ptr(), nwin, num_table, stream))
    return out


def resample(array: np.ndarray, rate_in: int, rate_out: int) -> np.ndarray:
    """numpy (frames, 2) in, numpy out in the input's dtype -- for the host route of checker.check."""
    from .engine import _require_cuda
    _require_cuda()
    device = torch.device("cuda", torch.cuda.current_device())
    a = np.ascontiguousarray(array)
    out = resample_on_device(torch.from_numpy(a.astype(np.float32, copy=False)).to(device), rate_in, rate_out)
    return out.cpu().numpy().astype(a.dtype if a.dtype in (np.float32, np.float64) else np.float64)
