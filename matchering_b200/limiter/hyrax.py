"""GPU drop-in for matchering/limiter/hyrax.py:78-99 `limit(array, config)`."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _native
from ..defaults import Config
from ..engine import _require_cuda, _stream_ptr, limiter_params, to_device_f32, to_host_like
from ..log import debug
from ..plan import limiter_constants


def limit(array, config: Config):
    """Returns the limited (N, 2) array; like the reference, returns `array` ITSELF (same object)
    when no frame exceeds the threshold (hyrax.py:83-85)."""
    _require_cuda()
    lib = _native.load()
    params = limiter_params(limiter_constants(config))
    device = torch.device("cuda", torch.cuda.current_device())
    debug("The limiter is started. Preparing the gain envelope...")
    with torch.cuda.device(device):
        x = to_device_f32(array, device)
        frames = x.shape[0]
        if frames <= 6:
            raise ValueError("The length of the input vector x must be greater than padlen, which is 6.")
        ws_bytes = int(lib.mgb_limiter_workspace_bytes(C.byref(params), frames))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        out = torch.empty_like(x)
        engaged = torch.zeros(1, dtype=torch.int32, device=device)
        _native.check(lib, lib.mgb_limit(C.byref(params), x.data_ptr(), out.data_ptr(), frames, ws.data_ptr(),
                                         ws_bytes, engaged.data_ptr(), _stream_ptr(device)))
        if int(engaged.item()) == 0:
            debug("The limiter is not needed!")
            return array
        return to_host_like(out, array)
