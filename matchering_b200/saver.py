"""Result writer (reference: matchering/saver.py:27-33)."""
from . import wavio
from .log import debug
from .results import real_soundfile


def save(file: str, result, sample_rate: int, subtype: str, name: str = "result") -> None:
    debug(f"Saving the {name.upper()} {sample_rate} Hz Stereo {subtype} to: '{file}'...")
    sf = real_soundfile()
    if sf is not None:
        sf.write(file, result, sample_rate, subtype)
    else:
        wavio.write(file, result, sample_rate, subtype)
    debug(f"'{file}' is saved")
