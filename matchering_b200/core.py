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
    temp_folder = config.temp_folder if config.temp_folder else get_temp_folder(results)

    target_audio, target_rate = _load_and_check(target, "target", temp_folder, config)
    reference_audio, reference_rate = _load_and_check(reference, "reference", temp_folder, config)
    if not config.allow_equality:
        _check_equality_any(target_audio, reference_audio)
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
    if preview_target or preview_result:
        from .preview_creator import create_preview
        first = next(item for item in (limited, plain, normalized) if item is not None)
        create_preview(target_audio, first, config, preview_target, preview_result)
    debug_line()
    info(Code.INFO_COMPLETED)


def _export(wanted: Result, audio, sample_rate: int, name: str = "result") -> None:
    """16/24-bit WAV results are quantised on the device and written as they come back (a quarter of
    the device->host bytes of a float64 array); everything else takes the reference's route through
    a host float array and `save`."""
    from .results import real_soundfile
    is_wav = wanted.file.lower().endswith(".wav")
    if is_wav and wanted.subtype in ("PCM_16", "PCM_24") and real_soundfile() is None:
        from . import wavio
        from .engine import encode_pcm
        bits = int(wanted.subtype[4:])
        debug(f"Saving the {name.upper()} {sample_rate} Hz Stereo {wanted.subtype} to: '{wanted.file}'...")
        wavio.write_pcm(wanted.file, encode_pcm(audio, bits), sample_rate, bits)
        debug(f"'{wanted.file}' is saved")
    else:
        save(wanted.file, audio.cpu().numpy().astype("float64"), sample_rate, wanted.subtype, name)


def _load_and_check(file: str, name: str, temp_folder: str, config: Config):
    """16/24-bit PCM WAV at the internal rate is decoded and checked on the device; everything else
    goes through the host loader and checker like in the reference."""
    from .device_io import check_on_device, load_to_device
    on_device = load_to_device(file, name, config)
    if on_device is not None:
        audio, rate = on_device
        return check_on_device(audio, config, name, rate), config.internal_sample_rate
    audio, rate = load(file, name, temp_folder)
    return check(audio, rate, config, name)


def _check_equality_any(target_audio, reference_audio) -> None:
    import torch
    if isinstance(target_audio, torch.Tensor) and isinstance(reference_audio, torch.Tensor):
        from .device_io import check_equality_on_device
        check_equality_on_device(target_audio, reference_audio)
    else:
        as_np = lambda a: a.cpu().numpy() if isinstance(a, torch.Tensor) else a
        check_equality(as_np(target_audio), as_np(reference_audio))
