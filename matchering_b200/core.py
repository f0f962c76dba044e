"""mg.process (reference: matchering/core.py:32-121): files in, mastered files out.  Loading,
validation and saving are host plumbing; `stages.main` in the middle runs on the GPU."""
from .checker import check, check_equality
from .defaults import Config
from .loader import load
from .log import Code, ModuleError, debug, debug_line, info
from .results import Result
from .saver import save
from .stages import main_device
from .utils import get_temp_folder

_DEFAULT_CONFIG = None


def process(target: str, reference: str, results: list, config: Config = None,
            preview_target: Result = None, preview_result: Result = None):
    global _DEFAULT_CONFIG
    if config is None:
        if _DEFAULT_CONFIG is None:
            _DEFAULT_CONFIG = Config()
        config = _DEFAULT_CONFIG
    debug_line()
    info(Code.INFO_LOADING)
    if not results:
        raise RuntimeError("The result list is empty")
    if preview_target or preview_result:
        raise NotImplementedError("previews are outside this build's scope (SURVEY.md section 8f)")
    temp_folder = config.temp_folder if config.temp_folder else get_temp_folder(results)

    target_audio, target_rate = load(target, "target", temp_folder)
    target_audio, target_rate = check(target_audio, target_rate, config, "target")
    reference_audio, reference_rate = load(reference, "reference", temp_folder)
    reference_audio, reference_rate = check(reference_audio, reference_rate, config, "reference")
    if not config.allow_equality:
        check_equality(target_audio, reference_audio)
    if (not (target_rate == reference_rate == config.internal_sample_rate)
            or not (target_audio.shape[1] == reference_audio.shape[1] == 2)
            or not (target_audio.shape[0] > config.fft_size and reference_audio.shape[0] > config.fft_size)):
        raise ModuleError(Code.ERROR_VALIDATION)

    limited, plain, normalized = main_device(
        target_audio, reference_audio, config,
        need_default=any(r.use_limiter for r in results),
        need_no_limiter=any(not r.use_limiter and not r.normalize for r in results),
        need_no_limiter_normalized=any(not r.use_limiter and r.normalize for r in results))

    debug_line()
    info(Code.INFO_EXPORTING)
    for wanted in results:
        audio = limited if wanted.use_limiter else (normalized if wanted.normalize else plain)
        _export(wanted, audio, config.internal_sample_rate)
    debug_line()
    info(Code.INFO_COMPLETED)


def _export(wanted: Result, audio, sample_rate: int) -> None:
    """16/24-bit WAV results are quantised on the device and written as they come back (a quarter of
    the device->host bytes of a float64 array); everything else takes the reference's route through
    a host float array and `save`."""
    from .results import real_soundfile
    is_wav = wanted.file.lower().endswith(".wav")
    if is_wav and wanted.subtype in ("PCM_16", "PCM_24") and real_soundfile() is None:
        from . import wavio
        from .engine import encode_pcm
        bits = int(wanted.subtype[4:])
        debug(f"Saving the RESULT {sample_rate} Hz Stereo {wanted.subtype} to: '{wanted.file}'...")
        wavio.write_pcm(wanted.file, encode_pcm(audio, bits), sample_rate, bits)
        debug(f"'{wanted.file}' is saved")
    else:
        save(wanted.file, audio.cpu().numpy().astype("float64"), sample_rate, wanted.subtype)
