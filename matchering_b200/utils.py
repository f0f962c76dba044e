"""Small host-side helpers with the names the reference exposes (matchering/utils.py:28-59): callers of
the reference package import them from here, so names, arguments and results are kept."""
import datetime
import math
import os
import random
import string

_ALPHABET = string.ascii_lowercase + string.digits


def get_temp_folder(results: list) -> str:
    """Directory of the first requested result: where temporary decodes go (utils.py:28-30)."""
    return os.path.dirname(os.path.abspath(results[0].file))


def random_str(size: int = 16) -> str:
    """`size` characters out of [a-z0-9] (utils.py:33-34)."""
    return "".join(random.choice(_ALPHABET) for _ in range(size))


def random_file(prefix: str = "", extension: str = "wav") -> str:
    """"<prefix>-<16 random characters>.<extension>", without the dash when there is no prefix (utils.py:37-39)."""
    stem = random_str()
    if prefix:
        stem = prefix + "-" + stem
    return stem + "." + extension


def to_db(value: float) -> str:
    """Amplitude ratio as decibels with four decimals, e.g. "-6.0206 dB" (utils.py:42-47)."""
    decibels = 20.0 * math.log10(value)
    return "%.4f dB" % decibels


def ms_to_samples(value: float, sample_rate: int) -> int:
    """Milliseconds -> whole samples, truncated (utils.py:50-51): limiter window lengths come from here."""
    return int(sample_rate * value * 1e-3)


def make_odd(value: int) -> int:
    """The next odd integer at or above `value` (utils.py:54-55): sliding-window sizes must be odd."""
    return value | 1


def time_str(length, sample_rate) -> str:
    """Whole seconds of `length` samples as H:MM:SS (utils.py:58-59)."""
    return str(datetime.timedelta(seconds=length // sample_rate))
