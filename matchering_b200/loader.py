"""mg.load (reference: matchering/loader.py:30-74): file -> ((frames, channels) float64, rate)."""
import os
import subprocess

from . import wavio
from .log import Code, ModuleError, debug, info, warning
from .results import real_soundfile
from .utils import random_file


def _read(file: str):
    sf = real_soundfile()
    if sf is not None:
        return sf.read(file, always_2d=True)
    return wavio.read(file)


def load(file: str, file_type: str, temp_folder: str):
    file_type = file_type.upper()
    debug(f"Loading the {file_type} file: '{file}'...")
    sound = rate = None
    try:
        sound, rate = _read(file)
    except (RuntimeError, OSError) as e:
        debug(str(e))
        if "unknown format" in str(e) or "Format not recognised" in str(e):
            sound, rate = _via_ffmpeg(file, file_type, temp_folder)
    if sound is None or rate is None:
        raise ModuleError(Code.ERROR_TARGET_LOADING if file_type == "TARGET" else Code.ERROR_REFERENCE_LOADING)
    debug(f"The {file_type} file is loaded")
    return sound, rate


def _via_ffmpeg(file: str, file_type: str, temp_folder: str):
    debug(f"Trying to load '{file}' with ffmpeg...")
    temp = os.path.join(temp_folder, random_file(prefix="temp"))
    try:
        with open(os.devnull, "w") as sink:
            subprocess.check_call(["ffmpeg", "-i", file, temp], stdout=sink, stderr=sink)
        sound, rate = _read(temp)
        os.remove(temp)
    except FileNotFoundError:
        debug("ffmpeg is not found in the system! Download, install and add it to PATH")
        return None, None
    except (subprocess.CalledProcessError, RuntimeError):
        debug(f"ffmpeg cannot convert '{file}' to .wav!")
        return None, None
    if file_type == "TARGET":
        warning(Code.WARNING_TARGET_IS_LOSSY)
    else:
        info(Code.INFO_REFERENCE_IS_LOSSY)
    return sound, rate
