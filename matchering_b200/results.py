"""Result descriptors (reference: matchering/results.py:25-46).  The reference validates
(extension, subtype) through libsndfile; this build writes WAV itself (saver.py) and accepts the
WAV subtypes it can encode -- anything else raises TypeError exactly where the reference would
for a format libsndfile does not know."""
import os

WAV_SUBTYPES = ("PCM_16", "PCM_24", "PCM_32", "FLOAT", "DOUBLE")
_FORMATS = {"WAV": WAV_SUBTYPES}


def real_soundfile():
    """libsndfile's binding when it is really installed (not a test stand-in), else None."""
    try:
        import soundfile as sf
    except ImportError:
        return None
    return sf if getattr(sf, "__libsndfile_version__", None) else None


def _check_format(ext: str, subtype: str = None) -> bool:
    sf = real_soundfile()  # used when present, so FLAC/AIFF/... keep working
    if sf is not None:
        return bool(sf.check_format(ext, subtype))
    if ext not in _FORMATS:
        return False
    return subtype is None or subtype in _FORMATS[ext]


class Result:
    def __init__(self, file: str, subtype: str, use_limiter: bool = True, normalize: bool = True):
        ext = os.path.splitext(file)[1][1:].upper()
        if not _check_format(ext):
            raise TypeError(f"{ext} format is not supported")
        if not _check_format(ext, subtype):
            raise TypeError(f"{ext} format does not have {subtype} subtype")
        self.file = file
        self.subtype = subtype
        self.use_limiter = use_limiter
        self.normalize = normalize


def pcm16(file: str) -> Result:
    return Result(file, "PCM_16")


def pcm24(file: str) -> Result:
    return Result(file, "PCM_24")
