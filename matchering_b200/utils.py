import math
import os
import random
import string
from datetime import timedelta


def get_temp_folder(results: list) -> str:
    return os.path.dirname(os.path.abspath(results[0].file))


def random_str(size: int = 16) -> str:
    return "".join(random.choices(string.ascii_lowercase + string.digits, k=size))


def random_file(prefix: str = "", extension: str = "wav") -> str:
    head = f"{prefix}-" if prefix else ""
    return f"{head}{random_str()}.{extension}"


def to_db(value: float) -> str:
    return f"{20 * math.log10(value):.4f} dB"


def ms_to_samples(value: float, sample_rate: int) -> int:
    return int(sample_rate * value * 1e-3)


def make_odd(value: int) -> int:
    return value if value & 1 else value + 1


def time_str(length, sample_rate) -> str:
    return str(timedelta(seconds=length // sample_rate))
