"""Preview creator (reference: matchering/preview_creator.py:30-94): the window of
`config.preview_size` frames in which the RESULT is loudest, cut from the result and from the target
clipped at the threshold, both faded in and out, saved as two files.

On the device: the per-window energies are one reduction kernel (mgb_window_energy), the argmax of a
few dozen numbers happens on the host, and the two pieces are cut / clipped / faded by
mgb_preview_piece; 16/24-bit WAV previews are quantised on the device like the results are.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native
from .defaults import Config
from .engine import _require_cuda, _stream_ptr, to_device_f32
from .log import Code, debug, debug_line, info
from .results import Result
from .utils import time_str


def loudest_window(result: torch.Tensor, config: Config):
    """-> (index, first frame, frames) of the preview window (dsp.strided_app_2d + batch_rms_2d + argmax)."""
    lib = _native.load()
    frames = int(result.shape[0])
    size, step = int(config.preview_size), int(config.preview_analysis_step)
    if size > frames:
        return 0, 0, frames
    count = (frames - size) // step + 1
    energy = torch.empty(count, dtype=torch.float64, device=result.device)
    _native.check(lib, lib.mgb_window_energy(result.data_ptr(), frames, size, step, count, energy.data_ptr(),
                                             _stream_ptr(result.device)))
    rms = np.sqrt(energy.cpu().numpy() / (2 * size))
    index = int(np.argmax(rms))
    return index, index * step, size


def cut_piece(signal: torch.Tensor, first: int, frames: int, clip_to: float, fade: int) -> torch.Tensor:
    lib = _native.load()
    out = torch.empty((frames, 2), dtype=torch.float32, device=signal.device)
    _native.check(lib, lib.mgb_preview_piece(signal[first:].data_ptr(), out.data_ptr(), frames, float(clip_to), int(fade),
                                             _stream_ptr(signal.device)))
    return out


def preview_pieces(target, result, config: Config):
    """-> (window index, target piece, result piece) as float32 CUDA tensors."""
    _require_cuda()
    device = result.device if isinstance(result, torch.Tensor) and result.is_cuda else torch.device(
        "cuda", torch.cuda.current_device())
    with torch.cuda.device(device):
        target = to_device_f32(target, device)
        result = to_device_f32(result, device)
        if target.shape != result.shape:
            raise ValueError("the preview needs the target and the result to be equally long")
        index, first, frames = loudest_window(result, config)
        fade = 0
        if frames != result.shape[0]:
            fade = int(min(config.preview_fade_size, frames // config.preview_fade_coefficient))
        return (index, cut_piece(target, first, frames, config.threshold, fade),
                cut_piece(result, first, frames, 0.0, fade))


def create_preview(target, result, config: Config, preview_target: Result, preview_result: Result) -> None:
    from .core import _export
    debug_line()
    info(Code.INFO_MAKING_PREVIEWS)
    sr = config.internal_sample_rate
    debug(f"The maximum duration of the preview is {config.preview_size / sr} seconds, "
          f"with the analysis step of {config.preview_analysis_step / sr} seconds")
    index, target_piece, result_piece = preview_pieces(target, result, config)
    begin = int(config.preview_analysis_step) * index
    debug(f"The best part to preview: {time_str(begin, sr)} - {time_str(begin + result_piece.shape[0], sr)}")
    if preview_target:
        _export(preview_target, target_piece, sr, "target preview")
    if preview_result:
        _export(preview_result, result_piece, sr, "result preview")
