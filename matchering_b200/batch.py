"""Batches of tracks with host buffers: `depth` tracks in flight per GPU (mgb_pipeline_*).

A single `stages.main` call is bound by its own PCIe copies (190 MB for a 3-minute track against
well under a millisecond of kernels); independent tracks overlap them: while track k computes, track
k+1 is already arriving and track k-1's result is leaving.  This is the entry point for the
"batch of tracks sharded one-per-GPU" configuration; across GPUs use one process per GPU and
`sharding.tracks_for_rank`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native
from .engine import get_plan


class MasteringPipeline:
    def __init__(self, config, max_target_frames: int, max_reference_frames: int, depth: int = 3, device=None):
        self.plan = get_plan(config, device)
        self.lib = self.plan.lib
        self.device = self.plan.device
        self.depth = depth
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _native.check(self.lib, self.lib.mgb_pipeline_create(C.byref(self.plan.struct), max_target_frames,
                                                                 max_reference_frames, depth, C.byref(self._handle)))
        self._pending = {}  # slot -> (target, reference, out) kept alive until the slot is waited on

    @staticmethod
    def _ptr(a):
        if isinstance(a, torch.Tensor):
            assert a.dtype == torch.float32 and a.is_contiguous() and not a.is_cuda
            return a.data_ptr(), a.shape[0]
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data, a.shape[0]

    def submit(self, target, reference, out) -> int:
        """target, reference, out: (frames, 2) float32 host arrays (numpy or torch, ideally pinned).
        Returns the slot to `wait` on; `out` is valid once that wait returns."""
        tp, tn = self._ptr(target)
        rp, rn = self._ptr(reference)
        op, on = self._ptr(out)
        assert on == tn
        slot = C.c_int32()
        with torch.cuda.device(self.device):
            _native.check(self.lib, self.lib.mgb_pipeline_submit(self._handle, tp, tn, rp, rn, op, C.byref(slot)))
        self._pending[slot.value] = (target, reference, out)
        return slot.value

    def submit_pcm(self, target, reference, out) -> int:
        """PCM host buffers (what audio files hold): int16 arrays of shape (frames, 2), or uint8
        arrays of shape (frames, 6) for packed little-endian 24-bit.  `out` chooses the result's
        width the same way.  A quarter to a half of the PCIe bytes of `submit`."""
        def describe(a):
            a_np = a.numpy() if isinstance(a, torch.Tensor) else a
            if a_np.dtype == np.int16 and a_np.ndim == 2 and a_np.shape[1] == 2:
                bits = 16
            elif a_np.dtype == np.uint8 and a_np.ndim == 2 and a_np.shape[1] == 6:
                bits = 24
            else:
                raise TypeError("PCM buffers are int16 (frames, 2) or uint8 (frames, 6)")
            assert a_np.flags["C_CONTIGUOUS"]
            return a_np.ctypes.data, bits, a_np.shape[0]
        tp, tb, tn = describe(target)
        rp, rb, rn = describe(reference)
        op, ob, on = describe(out)
        assert on == tn
        slot = C.c_int32()
        with torch.cuda.device(self.device):
            _native.check(self.lib, self.lib.mgb_pipeline_submit_pcm(self._handle, tp, tb, tn, rp, rb, rn, op, ob,
                                                                     C.byref(slot)))
        self._pending[slot.value] = (target, reference, out)
        return slot.value

    def wait(self, slot: int) -> _native.TrackState:
        st = _native.TrackState()
        _native.check(self.lib, self.lib.mgb_pipeline_wait(self._handle, slot, C.byref(st)))
        self._pending.pop(slot, None)
        return st

    def wait_all(self) -> None:
        for slot in list(self._pending):
            self.wait(slot)

    def streams(self):
        h, c, d = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _native.check(self.lib, self.lib.mgb_pipeline_streams(self._handle, C.byref(h), C.byref(c), C.byref(d)))
        return h.value, c.value, d.value

    def close(self) -> None:
        if self._handle:
            self.wait_all()
            self.lib.mgb_pipeline_destroy(self._handle)
            self._handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def master_many(pairs, config, depth: int = 3):
    """[(target, reference), ...] float32 host arrays -> list of limited results (numpy float32)."""
    pairs = list(pairs)
    if not pairs:
        return []
    max_t = max(len(t) for t, _ in pairs)
    max_r = max(len(r) for _, r in pairs)
    outs = [np.empty((len(t), 2), dtype=np.float32) for t, _ in pairs]
    with MasteringPipeline(config, max_t, max_r, depth) as pipe:
        for (t, r), o in zip(pairs, outs):
            pipe.submit(np.ascontiguousarray(t, dtype=np.float32), np.ascontiguousarray(r, dtype=np.float32), o)
        pipe.wait_all()
    return outs
