"""mg.load (reference: matchering/loader.py:30-74): file -> ((frames, channels) float64, rate).  Files libsndfile /
the built-in WAV reader cannot decode raise the reference's loading error; the reference's ffmpeg conversion
of lossy formats (loader.py:44-71) is outside this package's scope."""
from . import wavio
from .log import Code, ModuleError, debug
from .results import real_soundfile


def _read(file: str):
    sf = real_soundfile()
    if sf is not None:
        return sf.read(file, always_2d=True)
    return wavio.read(file)


def load(file: str, file_type: str, temp_folder: str):
    file_type = file_type.upper()
    debug(f"Loading the {file_type} file: '{file}'...")
    try:
        sound, rate = _read(file)
    except (RuntimeError, OSError) as e:
        debug(str(e))
        raise ModuleError(Code.ERROR_TARGET_LOADING if file_type == "TARGET" else Code.ERROR_REFERENCE_LOADING)
    debug(f"The {file_type} file is loaded")
    return sound, rate
