"""English text for every Code (reference: matchering/log/explanations.py:37-71)."""
from .codes import Code

_TARGET, _REFERENCE = "TARGET", "REFERENCE"


def _side_messages(side: str) -> dict:
    return {
        "LOADING": f"Audio stream error in the {side} file",
        "EXCEEDED": f"Track length is exceeded in the {side} file",
        "TOO_SMALL": f"The track length is too small in the {side} file",
        "CHANNELS": f"The number of channels exceeded in the {side} file",
    }


_T, _R = _side_messages(_TARGET), _side_messages(_REFERENCE)

TEXT = {
    Code.INFO_UPLOADING: "Uploading files",
    Code.INFO_WAITING: "Queued for processing",
    Code.INFO_LOADING: "Loading and analysis",
    Code.INFO_MATCHING_LEVELS: "Matching levels",
    Code.INFO_MATCHING_FREQS: "Matching frequencies",
    Code.INFO_CORRECTING_LEVELS: "Correcting levels",
    Code.INFO_FINALIZING: "Final processing and saving",
    Code.INFO_EXPORTING: "Exporting various audio formats",
    Code.INFO_MAKING_PREVIEWS: "Making previews",
    Code.INFO_COMPLETED: "The task is completed",
    Code.INFO_TARGET_IS_MONO: "The TARGET audio is mono. Converting it to stereo...",
    Code.INFO_REFERENCE_IS_MONO: "The REFERENCE audio is mono. Converting it to stereo...",
    Code.INFO_REFERENCE_IS_RESAMPLED: "The REFERENCE audio was resampled",
    Code.INFO_REFERENCE_IS_LOSSY: "Presumably the REFERENCE audio format is lossy",
    Code.WARNING_TARGET_IS_CLIPPING: (
        "Audio clipping is detected in the TARGET file. It is highly recommended to use the non-clipping version"),
    Code.WARNING_TARGET_LIMITER_IS_APPLIED: (
        "The applied limiter is detected in the TARGET file. "
        "It is highly recommended to use the version without a limiter"),
    Code.WARNING_TARGET_IS_RESAMPLED: (
        "The TARGET audio sample rate and internal sample rate were different. The TARGET audio was resampled"),
    Code.WARNING_TARGET_IS_LOSSY: (
        "Presumably the TARGET audio format is lossy. "
        "It is highly recommended to use lossless audio formats (WAV, FLAC, AIFF)"),
    Code.ERROR_TARGET_LOADING: _T["LOADING"],
    Code.ERROR_TARGET_LENGTH_IS_EXCEEDED: _T["EXCEEDED"],
    Code.ERROR_TARGET_LENGTH_IS_TOO_SMALL: _T["TOO_SMALL"],
    Code.ERROR_TARGET_NUM_OF_CHANNELS_IS_EXCEEDED: _T["CHANNELS"],
    Code.ERROR_TARGET_EQUALS_REFERENCE: (
        "The TARGET and REFERENCE files are the same. They must be different so that Matchering makes sense"),
    Code.ERROR_REFERENCE_LOADING: _R["LOADING"],
    Code.ERROR_REFERENCE_LENGTH_LENGTH_IS_EXCEEDED: _R["EXCEEDED"],
    Code.ERROR_REFERENCE_LENGTH_LENGTH_TOO_SMALL: _R["TOO_SMALL"],
    Code.ERROR_REFERENCE_NUM_OF_CHANNELS_IS_EXCEEDED: _R["CHANNELS"],
    Code.ERROR_UNKNOWN: "Unknown error",
    Code.ERROR_VALIDATION: "Validation failed! Please let the developers know about this error!",
}


def explain(code: Code, show_code: bool = False) -> str:
    text = TEXT[code]
    return f"{code}: {text}" if show_code else text


def get_explanation_handler(show_codes: bool = False):
    return lambda code: explain(code, show_codes)
