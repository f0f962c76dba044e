"""mg.check (reference: matchering/checker.py:30-142): input validation in front of the hot path.
Host-side numpy; SURVEY.md section 8(f) lists it as a later candidate for the device."""
import numpy as np

from .defaults import Config
from .log import Code, ModuleError, debug, info, warning
from .utils import time_str


def _resample(array, old_rate: int, new_rate: int):
    try:
        import resampy
    except ImportError:
        resampy = None
    if resampy is not None and getattr(resampy, "__version__", None):  # a real install, not a test stand-in
        return resampy.resample(array, old_rate, new_rate, axis=0)
    # resampy absent: polyphase Kaiser-windowed sinc from scipy instead
    from math import gcd
    from scipy.signal import resample_poly
    g = gcd(int(old_rate), int(new_rate))
    return resample_poly(array, new_rate // g, old_rate // g, axis=0)


def _count_max_peaks(array):
    peak = np.abs(array).max()
    hits = np.count_nonzero(np.isclose(np.abs(array), peak))
    return peak, hits


def check(array: np.ndarray, sample_rate: int, config: Config, name: str):
    name = name.upper()
    is_target = name == "TARGET"
    frames, channels = array.shape[0], array.shape[1]
    debug(f"{name} audio length: {frames} samples ({time_str(frames, sample_rate)})")
    if sample_rate != config.internal_sample_rate:
        debug(f"Resampling {name} audio from {sample_rate} Hz to {config.internal_sample_rate} Hz...")
        array = _resample(array, sample_rate, config.internal_sample_rate)
        sample_rate = config.internal_sample_rate
        frames = array.shape[0]
        (warning if is_target else info)(
            Code.WARNING_TARGET_IS_RESAMPLED if is_target else Code.INFO_REFERENCE_IS_RESAMPLED)
    if frames > config.max_length * sample_rate:
        raise ModuleError(Code.ERROR_TARGET_LENGTH_IS_EXCEEDED if is_target
                          else Code.ERROR_REFERENCE_LENGTH_LENGTH_IS_EXCEEDED)
    if frames < config.fft_size:
        raise ModuleError(Code.ERROR_TARGET_LENGTH_IS_TOO_SMALL if is_target
                          else Code.ERROR_REFERENCE_LENGTH_LENGTH_TOO_SMALL)
    if channels == 1:
        info(Code.INFO_TARGET_IS_MONO if is_target else Code.INFO_REFERENCE_IS_MONO)
        array = np.repeat(array, 2, axis=1)
    elif channels != 2:
        raise ModuleError(Code.ERROR_TARGET_NUM_OF_CHANNELS_IS_EXCEEDED if is_target
                          else Code.ERROR_REFERENCE_NUM_OF_CHANNELS_IS_EXCEEDED)
    if is_target:
        peak, hits = _count_max_peaks(array)
        if peak < 1.0 and hits > config.limited_samples_threshold:
            warning(Code.WARNING_TARGET_LIMITER_IS_APPLIED)
        elif peak >= 1.0 and hits > config.clipping_samples_threshold:
            warning(Code.WARNING_TARGET_IS_CLIPPING)
    return array, sample_rate


def check_equality(target: np.ndarray, reference: np.ndarray) -> None:
    if target.shape == reference.shape and np.allclose(target, reference):
        raise ModuleError(Code.ERROR_TARGET_EQUALS_REFERENCE)
