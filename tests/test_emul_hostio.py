"""The host transport (matchering_b200/csrc/hostio.cu: worker threads + pinned chunk ring) and the two
single-call host entries built on it, in the emulator build: the ring, the thread hand-shakes and the
conversions run exactly as on the device build, only the DMA is a memcpy."""
import ctypes as C

import numpy as np
import pytest

import port
from emul_harness import aligned, aligned_copy, emul_lib, get_emul_plan, limiter_params, ptr, run_pipeline
from matchering_b200 import _native, plan as plan_mod


@pytest.fixture(scope="module")
def lib():
    return emul_lib()


@pytest.fixture(scope="module")
def io(lib):
    h = C.c_void_p()
    # tiny chunks and a short ring so that a 100k-sample array wraps the ring dozens of times
    _native.check(lib, lib.mgb_host_io_create(3, 1024, 3, C.byref(h)))
    assert lib.mgb_host_io_threads(h) == 3
    yield h
    lib.mgb_host_io_destroy(h)


def _pinned(lib, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.mgb_host_alloc(n)
    assert p
    return np.ctypeslib.as_array((C.c_uint8 * n).from_address(p)).view(dtype).reshape(shape), p


@pytest.mark.parametrize("samples", [0, 1, 1023, 1024, 1025, 3 * 1024, 100003])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_upload_and_download_round_trip(lib, io, samples, dtype):
    rng = np.random.default_rng(samples)
    src = rng.standard_normal(samples).astype(dtype)
    dev = aligned((max(samples, 1),), np.float32)
    _native.check(lib, lib.mgb_host_upload(io, src.ctypes.data if samples else ptr(dev), src.dtype.itemsize, ptr(dev), samples, None))
    assert np.array_equal(dev[:samples], src.astype(np.float32))
    # pageable destination: float32 chunks through the ring, widened by the workers
    back = np.full(samples, 7, dtype=dtype)
    wide = aligned((max(samples, 1),), np.float64)
    _native.check(lib, lib.mgb_host_download(io, ptr(dev), back.ctypes.data if samples else ptr(wide), back.dtype.itemsize, samples, None, None))
    assert np.array_equal(back, src.astype(np.float32).astype(dtype))
    # pinned destination: widened "on the device", one copy
    if samples:
        pinned, p = _pinned(lib, (samples,), dtype)
        pinned[...] = 7
        _native.check(lib, lib.mgb_host_download(io, ptr(dev), p, pinned.dtype.itemsize, samples, ptr(wide), None))
        assert np.array_equal(pinned, src.astype(np.float32).astype(dtype))
        lib.mgb_host_free(p)


def test_repeated_transfers_reuse_the_ring(lib, io):
    rng = np.random.default_rng(1)
    dev = aligned((50000,), np.float32)
    for k in range(20):
        src = rng.standard_normal(50000 - 777 * k)
        _native.check(lib, lib.mgb_host_upload(io, src.ctypes.data, 8, ptr(dev), len(src), None))
        assert np.array_equal(dev[:len(src)], src.astype(np.float32))


@pytest.mark.parametrize("stores", [0, 1, 2])
@pytest.mark.parametrize("whole_chunks", [0, 1])
@pytest.mark.parametrize("download_ring", [0, 1, 2])
def test_transport_switches_do_not_change_the_bytes(lib, io, stores, whole_chunks, download_ring):
    """Streaming stores (256 / 512 bit), whole chunks per worker, the ring route for pinned float64 results:
    tuning switches of the transport (mgb_set_option), every combination moves the same bytes."""
    switches = (("host_streaming_stores", stores, 2), ("host_split_chunks", whole_chunks, 0),
                ("host_download_ring", download_ring, 1), ("host_prefetch", 4096 * whole_chunks, 8192))
    try:
        for name, value, _ in switches:
            _native.check(lib, lib.mgb_set_option(name.encode(), value))
        rng = np.random.default_rng(7)
        for samples in (5, 1024, 40011):
            a, b = rng.standard_normal(samples), rng.standard_normal(samples + 3)
            dev_a, dev_b = aligned((samples,), np.float32), aligned((samples + 3,), np.float32)
            _native.check(lib, lib.mgb_host_upload(io, a.ctypes.data, 8, ptr(dev_a), samples, None))
            _native.check(lib, lib.mgb_host_upload(io, b[1:].ctypes.data, 8, ptr(dev_b), samples + 2, None))  # odd alignment
            assert np.array_equal(dev_a, a.astype(np.float32))
            assert np.array_equal(dev_b[:samples + 2], b[1:].astype(np.float32))
            wide = aligned((samples,), np.float64)
            pinned, p = _pinned(lib, (samples + 1,), np.float64)
            _native.check(lib, lib.mgb_host_download(io, ptr(dev_a), p + 8, 8, samples, ptr(wide), None))  # 8-byte aligned only
            assert np.array_equal(pinned[1:], a.astype(np.float32).astype(np.float64))
            lib.mgb_host_free(p)
    finally:
        for name, _, default in switches:
            lib.mgb_set_option(name.encode(), default)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_stages_main_host_matches_staged_calls(lib, io, dtype):
    cfg = port.OracleConfig(fft_size=1024, max_piece_size=0.3)
    t, r = port.synth_target(30000, 1), port.synth_reference(28000, 2)
    want, st_want, _, _, L = run_pipeline(cfg, t, r, need=(True, True, True))
    ep = get_emul_plan(cfg, False)
    T, R = len(t), len(r)
    bufs = _native.HostBuffers()
    keep = dict(t=aligned((T, 2), np.float32), r=aligned((R, 2), np.float32), res=aligned((T, 2), np.float32),
                out=aligned((T, 2), np.float32), wide=aligned((T, 2), np.float64), ws=aligned((L.workspace_bytes,), np.uint8),
                st=aligned((C.sizeof(_native.TrackState),), np.uint8))
    bufs.d_target_lr, bufs.d_reference_lr, bufs.d_result_lr = ptr(keep["t"]).value, ptr(keep["r"]).value, ptr(keep["res"]).value
    bufs.d_out_lr, bufs.d_wide, bufs.d_workspace, bufs.d_state = (ptr(keep["out"]).value, ptr(keep["wide"]).value,
                                                                 ptr(keep["ws"]).value, ptr(keep["st"]).value)
    ht, hr = t.astype(dtype), r.astype(dtype)
    outs = [np.zeros((T, 2), dtype=dtype) for _ in range(2)]
    pinned, p = _pinned(lib, (T, 2), dtype)  # one output goes the direct (pinned) route
    state = _native.TrackState()
    _native.check(lib, lib.mgb_stages_main_host(io, C.byref(ep.struct), C.byref(L), ht.ctypes.data, hr.ctypes.data,
                                                ht.dtype.itemsize, outs[0].ctypes.data, p, outs[1].ctypes.data,
                                                ht.dtype.itemsize, C.byref(bufs), C.byref(state), None))
    assert np.array_equal(outs[0], want[0].astype(dtype))
    assert np.array_equal(pinned, want[1].astype(dtype))
    assert np.array_equal(outs[1], want[2].astype(dtype))
    assert state.steps_done == st_want.steps_done and state.limiter_engaged == 1
    assert state.rms_coefficient == st_want.rms_coefficient
    lib.mgb_host_free(p)


def test_limit_host_and_its_early_out(lib, io):
    cfg = port.OracleConfig()
    params = limiter_params(plan_mod.limiter_constants(cfg))
    n = 20000
    x = port.synth_limiter_input(n, 5)
    ws_bytes = int(lib.mgb_limiter_workspace_bytes(C.byref(params), n))
    d_in, d_out, wide = aligned((n, 2), np.float32), aligned((n, 2), np.float32), aligned((n, 2), np.float64)
    ws, flag = aligned((ws_bytes,), np.uint8), aligned((4,), np.int32)
    for scale, engaged_want in ((1.0, 1), (0.2, 0)):
        h_in = (scale * x).astype(np.float64)
        h_out = np.full((n, 2), 5.0)
        engaged = C.c_int32(-1)
        _native.check(lib, lib.mgb_limit_host(io, C.byref(params), h_in.ctypes.data, 8, h_out.ctypes.data, 8, n, ptr(d_in), ptr(d_out),
                                              ptr(wide), ptr(ws), ws_bytes, ptr(flag), C.byref(engaged), None))
        assert engaged.value == engaged_want
        if engaged_want:
            want = port.limit(h_in.astype(np.float32).astype(np.float64), cfg)
            assert np.abs(h_out - want).max() < 3e-7
        else:
            assert np.all(h_out == 5.0)  # untouched: the caller returns its input object
