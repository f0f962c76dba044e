"""The CUDA kernel sources, compiled for the host through tests/emul (a CUDA-on-CPU emulator),
checked against the oracle.  This validates kernel LOGIC in the build container where there is
no GPU; the same assertions run against the real library in test_gpu_parity.py (-m gpu)."""
import ctypes as C

import numpy as np
import pytest

import port
from emul_harness import EmulPlan, aligned, aligned_copy, emul_lib, limiter_params, ptr, run_pipeline
from matchering_b200 import _native, plan as plan_mod

TOL = 1e-5  # north-star bound on the sample-wise max-abs error of float32 results


@pytest.fixture(scope="module")
def lib():
    return emul_lib()


@pytest.mark.parametrize("n,f64", [(512, 0), (512, 1), (1024, 0), (4096, 0), (8192, 0), (16384, 0), (2048, 1), (8192, 1)])
def test_fft_matches_numpy(lib, n, f64):
    rng = np.random.default_rng(n + f64)
    dt = np.complex128 if f64 else np.complex64
    x = (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))).astype(dt)
    p = _native.Plan()
    p.fft_size = min(n, 8192) if n <= 8192 else n // 2
    p.n_lin, p.n_log, p.lowess_k, p.lowess_nfit = p.fft_size // 2 + 1, 10, 2, 2
    bufs = [aligned((1 << 20,), np.uint8) for _ in range(4)]
    p.d_tw_f32_F, p.d_tw_f32_2F, p.d_tw_f64_F, p.d_tw_f64_2F = [b.ctypes.data for b in bufs]
    _native.check(lib, lib.mgb_plan_fill_twiddles(C.byref(p), None))
    tw = bufs[1] if n > 8192 else (bufs[2] if f64 else bufs[0])
    xin, out = aligned_copy(x), aligned((2, n), dt)
    for direction in (1, -1):
        _native.check(lib, lib.mgb_test_fft(n, f64, direction, ptr(xin), ptr(out), 2, ptr(tw), None))
        want = np.fft.fft(x.astype(np.complex128), axis=1) if direction == 1 else np.fft.ifft(x.astype(np.complex128), axis=1) * n
        err = np.abs(out - want).max() / np.abs(want).max()
        assert err < (1e-14 if f64 else 5e-7)


def test_fir_design_matches_oracle(lib):
    cfg = port.OracleConfig()
    ep = EmulPlan(cfg)
    rng = np.random.default_rng(3)
    k = np.arange(2049)
    at = 1e-3 * (1 + 0.3 * rng.standard_normal(2049)) ** 2 + 1e-5
    ar = 3e-3 / np.sqrt(1 + k / 20.0) * (1 + 0.3 * rng.standard_normal(2049)) ** 2 + 1e-6
    avg = aligned_copy(np.stack([at, 0.5 * at, ar, 0.7 * ar]))
    fir = aligned((2, 4096), np.float64)
    ws = aligned((4 << 20,), np.uint8)
    _native.check(lib, lib.mgb_test_design_fir(C.byref(ep.struct), ptr(avg), ptr(fir), ptr(ws), None))
    assert np.abs(fir[0] - port.design_fir(at.copy(), ar.copy(), cfg)).max() < 1e-13
    assert np.abs(fir[1] - port.design_fir(0.5 * at, 0.7 * ar, cfg)).max() < 1e-13
    assert fir[0][0] == 0.0


def test_smoothing_operator_matches_direct_path_and_oracle(lib):
    """The Config-only smoothing matrix (built on the device from the direct kernels) against the
    direct chain and the oracle; a small grid keeps the emulated build of the matrix short."""
    cfg = port.OracleConfig(fft_size=1024, lin_log_oversampling=1, lowess_frac=0.06)
    ep = EmulPlan(cfg, operator=True)
    n = cfg.fft_size // 2 + 1
    rng = np.random.default_rng(8)
    at = 1e-3 * (1 + 0.3 * rng.standard_normal(n)) ** 2 + 1e-5
    ar = 2e-3 * (1 + 0.3 * rng.standard_normal(n)) ** 2 + 1e-6
    avg = aligned_copy(np.stack([at, 0.5 * at, ar, 0.7 * ar]))
    ws = aligned((4 << 20,), np.uint8)
    want = port.design_fir(at.copy(), ar.copy(), cfg)
    got = {}
    for direct in (0, 1):
        lib.mgb_set_option(b"design_direct", direct)
        fir = aligned((2, cfg.fft_size), np.float64)
        _native.check(lib, lib.mgb_test_design_fir(C.byref(ep.struct), ptr(avg), ptr(fir), ptr(ws), None))
        got[direct] = fir.copy()
    lib.mgb_set_option(b"design_direct", 0)
    assert np.abs(got[0][0] - want).max() < 1e-13 and np.abs(got[1][0] - want).max() < 1e-13
    assert np.abs(got[0] - got[1]).max() < 1e-13
    # and through the whole pipeline
    t, r = port.synth_target(30000, 1), port.synth_reference(28000, 2)
    cfg2 = port.OracleConfig(fft_size=1024, lin_log_oversampling=1, lowess_frac=0.06, max_piece_size=0.2)
    outs, _, _, _, _ = run_pipeline(cfg2, t, r, operator=True)
    _compare(outs, port.main(t.astype(np.float64), r.astype(np.float64), cfg2, True, True, True))


def _compare(outs, want):
    for got, ref in zip(outs, want):
        if got is None:
            assert ref is None
            continue
        assert np.abs(got - ref).max() < TOL


@pytest.mark.parametrize("tma", [1, 0])
def test_pipeline_matches_golden(lib, golden, tma):
    g = golden("pipeline_small.npz")
    cfg = port.OracleConfig(max_piece_size=float(g["max_piece_size_s"]))
    outs, st, fir, _, L = run_pipeline(cfg, g["target"], g["reference"], tma=tma)
    assert (L.target_divisions, L.target_piece) == (int(g["target_divisions"]), int(g["target_piece"]))
    assert (L.reference_divisions, L.reference_piece) == (int(g["reference_divisions"]), int(g["reference_piece"]))
    assert abs(st.rms_coefficient - float(g["rms_coefficient"])) < 1e-9
    assert abs(st.final_amplitude_coef - float(g["final_amplitude_coefficient"])) < 1e-12
    assert np.abs(fir[0] - g["fir_mid"]).max() < 1e-7 and np.abs(fir[1] - g["fir_side"]).max() < 1e-7
    _compare(outs, (g["limited"], g["no_limiter"], g["normalized"]))
    assert st.limiter_engaged == 1


def test_pipeline_quiet_reference_early_out(lib, golden):
    g = golden("pipeline_quiet_reference.npz")
    cfg = port.OracleConfig(max_piece_size=float(g["max_piece_size_s"]))
    outs, st, _, _, _ = run_pipeline(cfg, g["target"], g["reference"], need=(True, True, False))
    assert st.limiter_engaged == 0 and st.final_amplitude_coef < 1.0
    _compare(outs, (g["limited"], g["no_limiter"], None))


def test_generic_convolution_kernel_matches_golden(lib, golden):
    """fft_size 4096 normally takes the fused convolution kernel; the all-shared-memory one must agree."""
    g = golden("pipeline_small.npz")
    cfg = port.OracleConfig(max_piece_size=float(g["max_piece_size_s"]))
    lib.mgb_set_option(b"conv_fused", 0)
    try:
        outs, st, _, _, _ = run_pipeline(cfg, g["target"], g["reference"])
    finally:
        lib.mgb_set_option(b"conv_fused", 1)
    _compare(outs, (g["limited"], g["no_limiter"], g["normalized"]))


@pytest.mark.parametrize("fft_size,n,piece_s", [(4096, 60011, 0.5), (2048, 30011, 0.25)])
def test_long_frame_convolution_against_oracle(lib, fft_size, n, piece_s):
    """Overlap-save frames of 4 FIR lengths (3F outputs per 4F-point transform pair; default where pieces are at
    least 3F long) against the oracle and against the 2-FIR-length frames: ragged last frame, piece boundaries
    inside frames, samples beyond piece*divisions."""
    cfg = port.OracleConfig(fft_size=fft_size, max_piece_size=piece_s)
    t, r = port.synth_target(n, 3), port.synth_reference(n - 777, 4)
    want = port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True)
    got = {}
    try:
        for frame in (4, 2):
            assert lib.mgb_set_option(b"conv_frame", frame) == 0
            outs, st, fir, _, L = run_pipeline(cfg, t, r)
            assert L.target_piece >= 3 * fft_size
            _compare(outs, want)
            got[frame] = outs
    finally:
        lib.mgb_set_option(b"conv_frame", 4)
    assert np.abs(got[2][1] - got[4][1]).max() < 2e-6


def test_persistent_convolution_and_chained_analysis_twiddles(lib):
    """Defaults: conv_persistent (a CTA walks several frames; the next frame's bulk copy is issued into the frame
    buffer once the last inverse pass has gathered from it) and analyze_chain (twiddle powers built in registers).
    Switched off, the one-frame-per-CTA convolution gives the same bits and the table-read analysis the same
    spectra to float32 rounding."""
    cfg = port.OracleConfig(max_piece_size=1.5)
    n = 12288 * 19 + 4097  # 20 frames of 3F outputs on the emulator's 8 SMs: two or three per CTA, ragged end
    t, r = port.synth_target(n, 5), port.synth_reference(n - 3001, 6)
    base, _, _, _, L = run_pipeline(cfg, t, r)
    assert L.target_piece >= 3 * 4096
    _compare(base, port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True))
    try:
        for name, tol in ((b"conv_persistent", 0.0), (b"analyze_chain", 1e-6)):
            assert lib.mgb_set_option(name, 0) == 0
            outs, _, _, _, _ = run_pipeline(cfg, t, r)
            assert lib.mgb_set_option(name, 1) == 0
            for a, b in zip(outs, base):
                assert np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max() <= tol, name
    finally:
        lib.mgb_set_option(b"conv_persistent", 1)
        lib.mgb_set_option(b"analyze_chain", 1)


@pytest.mark.parametrize("fft_size,sr", [(512, 44100), (512, 8000), (1024, 44100), (2048, 22050), (4096, 96000), (8192, 44100)])
def test_pipeline_other_configs(lib, fft_size, sr):
    cfg = port.OracleConfig(internal_sample_rate=sr, fft_size=fft_size, max_piece_size=0.4, rms_correction_steps=2)
    n = int(sr * 1.1) + 13
    t, r = port.synth_target(n, 5), port.synth_reference(n - 4001, 6)
    outs, st, _, _, _ = run_pipeline(cfg, t, r)
    want = port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True)
    _compare(outs, want)
    assert st.steps_done == 2


def test_pipeline_fft_size_16384(lib):
    """fft_size 16384: the analysis reads its frames straight from global memory, the design FFT's float64 planes and
    the convolution's 32768-point frames live in global memory (convolve_global_kernel: a CTA walks several frames,
    ragged last frame, pieces of two and a bit frames)."""
    cfg = port.OracleConfig(fft_size=16384, max_piece_size=1.0)
    n = 16384 * 20 + 777  # 21 frames on 16 CTAs: some walk two
    t, r = port.synth_target(n, 8), port.synth_reference(n - 5003, 9)
    outs, st, fir, _, L = run_pipeline(cfg, t, r)
    want = port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True)
    _compare(outs, want)
    assert st.steps_done == 4 and L.target_piece >= 16384


def _limit(lib, x, cfg):
    params = limiter_params(plan_mod.limiter_constants(cfg))
    n = len(x)
    ws_bytes = int(lib.mgb_limiter_workspace_bytes(C.byref(params), n))
    ws = aligned((ws_bytes,), np.uint8)
    xin, out = aligned_copy(x, np.float32), aligned((n, 2), np.float32)
    engaged = aligned((4,), np.int32)
    _native.check(lib, lib.mgb_limit(C.byref(params), ptr(xin), ptr(out), n, ptr(ws), ws_bytes, ptr(engaged), None))
    return out, int(engaged[0])


def test_limiter_matches_golden_several_chunks(lib, golden):
    g = golden("limiter.npz")
    out, engaged = _limit(lib, g["x"], port.OracleConfig())
    assert engaged == 1
    assert np.abs(out - g["y_44100"]).max() < 3e-7
    lib.mgb_set_option(b"limiter_ticket", 1)  # chunks by atomic ticket instead of by block index
    try:
        ticketed, _ = _limit(lib, g["x"], port.OracleConfig())
    finally:
        lib.mgb_set_option(b"limiter_ticket", 0)
    assert np.array_equal(ticketed, out)
    out96, _ = _limit(lib, g["x"], port.OracleConfig(internal_sample_rate=96000))
    assert np.abs(out96 - g["y_96000"]).max() < 3e-7


@pytest.mark.parametrize("n", [7, 100, 4607, 4608, 4609, 9217])
def test_limiter_edge_lengths(lib, n):
    cfg = port.OracleConfig()
    x = port.synth_limiter_input(max(n, 64), seed=n)[:n]
    out, engaged = _limit(lib, x, cfg)
    want = port.limit(x.astype(np.float64), cfg)
    assert np.abs(out - want).max() < 3e-7


@pytest.mark.parametrize("sr,attack", [(4000, 1.0), (8000, 1.0), (44100, 0.1), (44100, 4.0)])
def test_limiter_narrow_and_wide_windows(lib, sr, attack):
    """Attack windows of 3..350 samples: narrow ones take the per-sample window evaluation, the others the
    shared-core one (limiter.cu, P2)."""
    cfg = port.OracleConfig(internal_sample_rate=sr, limiter=port.OracleLimiterConfig(attack=attack))
    x = port.synth_limiter_input(12000, seed=sr + int(10 * attack))
    out, engaged = _limit(lib, x, cfg)
    want = port.limit(x.astype(np.float64), cfg)
    assert engaged == 1 and np.abs(out - want).max() < 3e-7


@pytest.mark.parametrize("sr,attack,hold_order", [(192000, 1.0, 1), (176400, 1.0, 2), (96000, 2.0, 1), (96000, 3.0, 1), (192000, 2.0, 1),
                                                  (192000, 2.0, 2)])
def test_limiter_wide_halos(lib, sr, attack, hold_order):
    """The default limiter at 176.4 / 192 kHz and slow attacks: the attack filter's warm-up needs 4200 ... 8130 samples
    of halo around a 4608-sample chunk (19 ... 25 span samples per thread, one CTA per SM)."""
    lim = port.OracleLimiterConfig(attack=attack, hold_filter_order=hold_order)
    cfg = port.OracleConfig(internal_sample_rate=sr, limiter=lim)
    lc = plan_mod.limiter_constants(cfg)
    assert (2 * lc.warmup + lc.hold + 2 * lc.reach + 64 > 4096) == (sr * attack > 180000)  # (176.4 kHz fits the old span)
    x = port.synth_limiter_input(30011, seed=sr // 100 + int(10 * attack))
    out, engaged = _limit(lib, x, cfg)
    want = port.limit(x.astype(np.float64), cfg)
    assert engaged == 1 and np.abs(out - want).max() < 3e-7


@pytest.mark.parametrize("hold_order,release_order,sr", [(2, 1, 44100), (1, 2, 44100), (2, 2, 44100), (2, 2, 96000),
                                                         (2, 2, 8000)])
def test_limiter_higher_filter_orders(lib, hold_order, release_order, sr):
    """LimiterConfig.hold_filter_order / release_filter_order > 1 (legal, matchering/defaults.py:48-56): the hold
    and release low-passes become order-N Butterworth sections, run as blocked scans over the filter's state
    vector with matrix-valued carries across threads, warps and chunks (five chunks here, so the look-back
    carries states, too)."""
    # (a release of 20 ms instead of 3 s: in half a second of signal the default release filter never rises above
    # the hold filter, and its order would not show in the output)
    lim = port.OracleLimiterConfig(hold_filter_order=hold_order, release_filter_order=release_order, release=20.0)
    cfg = port.OracleConfig(internal_sample_rate=sr, limiter=lim)
    x = port.synth_limiter_input(22000, seed=10 * hold_order + release_order)
    x[9000:15000] *= 0.05  # a quiet stretch: the gain recovers along the release curve
    want = port.limit(x.astype(np.float64), cfg)
    out, engaged = _limit(lib, x, cfg)
    assert engaged == 1
    # not the order-1 result, and neither section alone explains the difference
    for other in (port.OracleLimiterConfig(release=20.0),
                  port.OracleLimiterConfig(hold_filter_order=hold_order, release=20.0),
                  port.OracleLimiterConfig(release_filter_order=release_order, release=20.0)):
        if (other.hold_filter_order, other.release_filter_order) != (hold_order, release_order):
            plain = port.limit(x.astype(np.float64), port.OracleConfig(internal_sample_rate=sr, limiter=other))
            assert np.abs(plain - want).max() > 1e-4
    assert np.abs(out - want).max() < 1e-6
    for inclusive in (0,):  # aggregates only: every look-back walks its whole window
        lib.mgb_set_option(b"lookback_inclusive", inclusive)
        try:
            again, _ = _limit(lib, x, cfg)
        finally:
            lib.mgb_set_option(b"lookback_inclusive", 1)
        assert np.abs(again - want).max() < 1e-6


def test_limiter_not_engaged_copies_input(lib):
    x = (0.2 * port.synth_limiter_input(6000, 2)).astype(np.float32)
    out, engaged = _limit(lib, x, port.OracleConfig())
    assert engaged == 0 and np.array_equal(out, x)


def test_limiter_rejects_too_short_input(lib):
    with pytest.raises(ValueError):
        _limit(lib, np.zeros((6, 2), np.float32), port.OracleConfig())


def test_unsupported_configs_fail_loudly():
    for fft_size in (256, 32768):
        with pytest.raises(plan_mod.UnsupportedConfig):
            plan_mod.build_tables(port.OracleConfig(fft_size=fft_size))
    with pytest.raises(plan_mod.UnsupportedConfig):
        plan_mod.build_tables(port.OracleConfig(lowess_it=9))
    lim = port.OracleLimiterConfig(hold_filter_order=3)
    with pytest.raises(plan_mod.UnsupportedConfig):
        plan_mod.build_tables(port.OracleConfig(limiter=lim))


def test_pipeline_entry_points_match_single_track_path(lib, golden):
    """mgb_pipeline_* (batch entry, several slots) gives the same samples as the stage calls."""
    g = golden("pipeline_small.npz")
    cfg = port.OracleConfig(max_piece_size=float(g["max_piece_size_s"]))
    from emul_harness import get_emul_plan
    ep = get_emul_plan(cfg)
    handle = C.c_void_p()
    _native.check(lib, lib.mgb_pipeline_create(C.byref(ep.struct), 40000, 40000, 2, C.byref(handle)))
    try:
        outs, slots = [], []
        for k in range(3):  # more submissions than slots: a slot gets reused
            t = aligned_copy(g["target"][: 40000 - 7 * k], np.float32)
            r = aligned_copy(g["reference"], np.float32)
            o = aligned((len(t), 2), np.float32)
            slot = C.c_int32()
            _native.check(lib, lib.mgb_pipeline_submit(handle, ptr(t), len(t), ptr(r), len(r), ptr(o), C.byref(slot)))
            outs.append(o)
            slots.append(slot.value)
        assert slots == [0, 1, 0]
        st = _native.TrackState()
        _native.check(lib, lib.mgb_pipeline_wait(handle, 1, C.byref(st)))
        assert st.steps_done == 4
        assert np.abs(outs[0] - g["limited"]).max() < TOL
        want = port.main(g["target"][: 40000 - 7].astype(np.float64), g["reference"].astype(np.float64), cfg)[0]
        assert np.abs(outs[1] - want).max() < TOL
    finally:
        lib.mgb_pipeline_destroy(handle)


def _to_pcm24(x):
    q = np.clip(np.rint(x.astype(np.float64) * 8388607.0), -8388608, 8388607).astype(np.int64) & 0xFFFFFF
    return np.stack([q & 0xFF, (q >> 8) & 0xFF, (q >> 16) & 0xFF], axis=-1).astype(np.uint8).reshape(len(x), 6)


def _from_pcm24(b):
    b = b.reshape(-1, 3).astype(np.int64)
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    v = np.where(v >= 1 << 23, v - (1 << 24), v)
    return (v / 8388608.0).reshape(-1, 2)


def test_pcm_conversions_are_bit_exact(lib):
    rng = np.random.default_rng(4)
    x = np.clip(0.7 * rng.standard_normal((5001, 2)), -1.2, 1.2).astype(np.float32)
    xin = aligned_copy(x)
    q16 = aligned((5001, 2), np.int16)
    _native.check(lib, lib.mgb_pcm_encode(ptr(xin), 16, ptr(q16), x.size, None))
    want16 = np.clip(np.rint(x.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16)  # in float64, like wavio.write
    assert np.array_equal(q16, want16)
    q24 = aligned((5001, 6), np.uint8)
    _native.check(lib, lib.mgb_pcm_encode(ptr(xin), 24, ptr(q24), x.size, None))
    want24 = np.clip(np.rint(x.astype(np.float64) * 8388607.0), -8388608, 8388607).astype(np.int64)
    got24 = (_from_pcm24(q24) * 8388608.0).astype(np.int64)
    assert np.array_equal(got24, want24)
    back = aligned((5001, 2), np.float32)
    _native.check(lib, lib.mgb_pcm_decode(ptr(q16), 16, ptr(back), x.size, None))
    assert np.array_equal(back, q16.astype(np.float32) / np.float32(32768.0))
    _native.check(lib, lib.mgb_pcm_decode(ptr(q24), 24, ptr(back), x.size, None))
    assert np.array_equal(back, _from_pcm24(q24).astype(np.float32))


def test_pipeline_pcm_entry(lib, golden):
    """PCM16 in / PCM24 out through the batch entry == decode, float pipeline, encode."""
    g = golden("pipeline_small.npz")
    cfg = port.OracleConfig(max_piece_size=float(g["max_piece_size_s"]))
    from emul_harness import get_emul_plan
    ep = get_emul_plan(cfg)
    t16 = aligned_copy(np.clip(np.rint(g["target"] * 32767.0), -32768, 32767).astype(np.int16))
    r24 = aligned_copy(_to_pcm24(g["reference"]))
    out24 = aligned((len(t16), 6), np.uint8)
    handle = C.c_void_p()
    _native.check(lib, lib.mgb_pipeline_create(C.byref(ep.struct), 40000, 40000, 2, C.byref(handle)))
    try:
        slot = C.c_int32()
        _native.check(lib, lib.mgb_pipeline_submit_pcm(handle, ptr(t16), 16, len(t16), ptr(r24), 24, len(r24), ptr(out24), 24,
                                                       C.byref(slot)))
        _native.check(lib, lib.mgb_pipeline_wait(handle, slot.value, None))
    finally:
        lib.mgb_pipeline_destroy(handle)
    t = t16.astype(np.float64) / 32768.0
    r = _from_pcm24(r24)
    want = port.main(t, r, cfg)[0]
    got = _from_pcm24(out24) * (8388608.0 / 8388607.0)
    assert np.abs(got - want).max() < TOL + 1.0 / 8388607


@pytest.mark.parametrize("steps", [0, 1, 3])
def test_rms_correction_step_counts(lib, steps):
    cfg = port.OracleConfig(fft_size=1024, max_piece_size=0.3, rms_correction_steps=steps)
    t, r = port.synth_target(40000, 31), port.synth_reference(35000, 32)
    outs, st, _, _, _ = run_pipeline(cfg, t, r)
    _compare(outs, port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True))
    assert st.steps_done == steps


def test_shortest_legal_tracks_single_piece(lib):
    """Just over fft_size: one piece, one STFT frame, two convolution frames, one limiter chunk."""
    cfg = port.OracleConfig(fft_size=1024)
    t, r = port.synth_target(1500, 41), port.synth_reference(1025, 42)
    outs, st, _, _, L = run_pipeline(cfg, t, r)
    assert (L.target_divisions, L.target_piece, L.reference_divisions) == (1, 1500, 1)
    assert st.target_loud_pieces == 1 and st.reference_loud_pieces == 1  # sqrt(r*r) >= r
    _compare(outs, port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True))


def test_very_quiet_target_and_clipping_reference(lib):
    """Large level-matching gain (target 80 dB down) and a reference that peaks above the threshold
    (normalize leaves it alone, dsp.py:98)."""
    cfg = port.OracleConfig(fft_size=1024, max_piece_size=0.4)
    t = (1e-4 * port.synth_target(30000, 51)).astype(np.float32)
    r = np.clip(1.5 * port.synth_reference(30000, 52), -1.0, 1.0).astype(np.float32)
    outs, st, _, _, _ = run_pipeline(cfg, t, r)
    want = port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True)
    assert st.final_amplitude_coef == 1.0 and st.rms_coefficient > 1000
    _compare(outs, want)


def test_too_short_pieces_are_rejected(lib):
    from emul_harness import get_emul_plan
    cfg = port.OracleConfig(fft_size=4096, max_piece_size=0.1)  # 4410-sample pieces at most
    ep = get_emul_plan(cfg)
    L = _native.TrackLayout()
    rc = lib.mgb_track_layout_init(C.byref(ep.struct), 5000, 5000, C.byref(L))  # 2 pieces of 2500 < fft_size
    assert rc == _native.MGB_ERR_UNSUPPORTED


def test_limiter_lookback_through_aggregates_only(lib, golden):
    """With inclusive states withheld every chunk's carry is rebuilt from its predecessors'
    zero-carry aggregates, weighted by the pole's per-chunk decay, down to the 1e-9 cut-off (on a GPU
    this is what a chunk sees while its predecessors are still in flight)."""
    x = port.synth_limiter_input(4608 * 40 + 123, seed=9)
    want = port.limit(x.astype(np.float64), port.OracleConfig())
    lib.mgb_set_option(b"lookback_inclusive", 0)
    try:
        out, _ = _limit(lib, x, port.OracleConfig())
    finally:
        lib.mgb_set_option(b"lookback_inclusive", 1)
    assert np.abs(out - want).max() < 3e-7


def test_checker_reductions(lib):
    rng = np.random.default_rng(6)
    x = (0.5 * rng.standard_normal((20001, 2))).astype(np.float32)
    x[100, 0] = x[2000, 1] = -(np.abs(x).max() + 0.25)  # two samples at the (negative) peak
    xin = aligned_copy(x)
    scratch = aligned((16,), np.uint8)
    _native.check(lib, lib.mgb_check_peaks(ptr(xin), len(x), ptr(scratch), None))
    peak, hits = float(scratch[:4].view(np.float32)[0]), int(scratch[8:].view(np.uint64)[0])
    want_peak = np.abs(x).max()
    want_hits = np.count_nonzero(np.logical_or(np.isclose(x, want_peak), np.isclose(x, -want_peak)))  # dsp.py:49-54
    assert peak == want_peak and hits == want_hits == 2
    y = aligned_copy(x)
    diff = aligned((8,), np.uint8)
    _native.check(lib, lib.mgb_check_equality(ptr(xin), ptr(y), len(x), ptr(diff), None))
    assert int(diff.view(np.uint64)[0]) == 0
    y[5, 1] += 0.01
    _native.check(lib, lib.mgb_check_equality(ptr(xin), ptr(y), len(x), ptr(diff), None))
    assert int(diff.view(np.uint64)[0]) == 1


@pytest.mark.parametrize("side_level", [0.0, 1e-4])
def test_mono_and_near_mono_targets(lib, side_level):
    """A (nearly) mono target against a wide reference makes the side FIR ~1e4..1e6 times larger
    than the mid FIR.  mid and side share one complex transform per frame; without the per-frame
    power-of-two balancing the mid channel's rounding noise would swamp the side channel."""
    cfg = port.OracleConfig(fft_size=1024, max_piece_size=0.4)
    rng = np.random.default_rng(77)
    mid = port.synth_target(30000, 71)[:, :1]
    wobble = (side_level * rng.standard_normal((30000, 1))).astype(np.float32)
    t = np.ascontiguousarray(np.concatenate([mid + wobble, mid - wobble], axis=1))
    r = port.synth_reference(30000, 72)
    outs, st, _, _, _ = run_pipeline(cfg, t, r)
    assert st.fir_peak_side_bits > 30 * st.fir_peak_mid_bits
    want = port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True)
    _compare(outs, want)
    if side_level == 0.0:
        assert np.array_equal(outs[1][:, 0], outs[1][:, 1])  # stays exactly mono, like the reference


def test_preview_kernels_match_golden(lib, golden):
    """mgb_window_energy + mgb_preview_piece against the reference's create_preview (tests/golden/preview.npz)."""
    g = golden("preview.npz")
    sr, every = int(g["sample_rate"]), int(g["every"])
    size, step = int(g["preview_size_s"]) * sr, int(g["preview_analysis_step_s"]) * sr
    target, result = aligned_copy(g["target"]), aligned_copy(g["result"])
    n = len(result)
    count = (n - size) // step + 1
    energy = aligned((count,), np.float64)
    _native.check(lib, lib.mgb_window_energy(ptr(result), n, size, step, count, ptr(energy), None))
    want = np.array([np.sum(g["result"][w * step:w * step + size].astype(np.float64) ** 2) for w in range(count)])
    assert np.abs(energy - want).max() < 1e-9 * want.max()
    index = int(np.argmax(np.sqrt(energy / (2 * size))))
    assert index == int(g["index"])
    fade = min(1 * sr, size // 8)
    thr = port.OracleConfig().threshold
    for src, clip_to, piece, edge, where in ((target, thr, g["target_piece"], g["target_piece_head"], slice(0, 300)),
                                             (result, 0.0, g["result_piece"], g["result_piece_tail"], slice(-300, None))):
        out = aligned((size, 2), np.float32)
        _native.check(lib, lib.mgb_preview_piece(ptr(src[index * step:]), ptr(out), size, clip_to, fade, None))
        assert np.abs(out[::every] - piece).max() < 2e-7 and np.abs(out[where] - edge).max() < 2e-7
        assert out[0].tolist() == [0.0, 0.0] and out[-1].tolist() == [0.0, 0.0]
    # more windows than fit, or fades longer than the piece, are refused
    assert lib.mgb_window_energy(ptr(result), n, size, step, count + 1, ptr(energy), None) == _native.MGB_ERR_INVALID
    assert lib.mgb_preview_piece(ptr(result), ptr(out), 100, 0.0, 51, None) == _native.MGB_ERR_INVALID


@pytest.mark.parametrize("one_sided", [False, True])
def test_tonal_material_with_digital_silence(lib, one_sided):
    """Notes between stretches of exact zeros (silent pieces and frames), optionally with the right
    channel silent and a DC offset on the left."""
    cfg = port.OracleConfig(max_piece_size=0.3)
    n = 44100 * 2
    t = 0.5 * port.synth_tonal(n, 1)
    r = np.tanh(2.0 * port.synth_tonal(n + 500, 2)).astype(np.float32)
    if one_sided:
        t[:, 1] = 0.0
        t[:, 0] += 0.05
    outs, st, _, _, _ = run_pipeline(cfg, t, r)
    want = port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True)
    _compare(outs, want)


def test_finalize_twice_on_one_track(lib):
    """mgb_finalize is re-entrant on a mastered track: the second call must find fresh limiter
    tickets and look-back slots (it used to return MGB_OK and write nothing)."""
    from emul_harness import get_emul_plan
    cfg = port.OracleConfig(fft_size=1024, max_piece_size=0.3)
    t, r = port.synth_target(30000, 1), port.synth_reference(28000, 2)
    outs, state, fir, result, L = run_pipeline(cfg, t, r, need=(True, False, False))
    ep = get_emul_plan(cfg, False)
    # run_pipeline's workspace is gone; redo the stages on one we keep and finalize twice
    ws = aligned((L.workspace_bytes,), np.uint8)
    tgt, ref = aligned_copy(t, np.float32), aligned_copy(r, np.float32)
    res = aligned((len(t), 2), np.float32)
    st = _native.TrackState()
    P, LL = C.byref(ep.struct), C.byref(L)
    _native.check(lib, lib.mgb_match_levels(P, LL, ptr(tgt), ptr(ref), ptr(ws), C.byref(st), None))
    _native.check(lib, lib.mgb_match_frequencies(P, LL, ptr(tgt), ptr(res), None, ptr(ws), C.byref(st), None))
    _native.check(lib, lib.mgb_correct_levels(P, LL, ptr(ws), C.byref(st), None))
    first, second = aligned((len(t), 2), np.float32), aligned((len(t), 2), np.float32)
    second[...] = 7.0
    _native.check(lib, lib.mgb_finalize(P, LL, ptr(res), ptr(first), None, None, ptr(ws), C.byref(st), None))
    _native.check(lib, lib.mgb_finalize(P, LL, ptr(res), ptr(second), None, None, ptr(ws), C.byref(st), None))
    assert st.limiter_engaged == 1
    assert np.array_equal(first, second) and np.array_equal(first, outs[0])


@pytest.mark.parametrize("it", [1, 3])
def test_fir_design_with_lowess_robustness_iterations(lib, it):
    """Config.lowess_it > 0 (legal, matchering/defaults.py:135; dsp.py:103-106): LOWESS re-weights its
    regressions with bisquare weights of the residuals, the smoothing is no longer a Config-only matrix and
    the design kernel runs the chain directly.  A curve with outliers, so that the weights matter."""
    cfg = port.OracleConfig(lowess_it=it)
    ep = EmulPlan(cfg)
    assert ep.struct.lowess_it == it and not ep.struct.d_smooth_op
    rng = np.random.default_rng(30 + it)
    k = np.arange(2049)
    at = 1e-3 * (1 + 0.3 * rng.standard_normal(2049)) ** 2 + 1e-5
    ar = 3e-3 / np.sqrt(1 + k / 20.0) * (1 + 0.3 * rng.standard_normal(2049)) ** 2 + 1e-6
    ar[rng.choice(2049, 40, replace=False)] *= 30.0  # outliers: what the robustness iterations are for
    avg = aligned_copy(np.stack([at, 0.5 * at, ar, 0.7 * ar]))
    fir = aligned((2, 4096), np.float64)
    ws = aligned((4 << 20,), np.uint8)
    _native.check(lib, lib.mgb_test_design_fir(C.byref(ep.struct), ptr(avg), ptr(fir), ptr(ws), None))
    want = port.design_fir(at.copy(), ar.copy(), cfg)
    plain = port.design_fir(at.copy(), ar.copy(), port.OracleConfig())
    assert np.abs(want - plain).max() > 1e-4  # the iterations change the FIR visibly
    assert np.abs(fir[0] - want).max() < 1e-12
    assert np.abs(fir[1] - port.design_fir(0.5 * at, 0.7 * ar, cfg)).max() < 1e-12


def test_pipeline_with_lowess_robustness_iterations(lib):
    cfg = port.OracleConfig(fft_size=1024, max_piece_size=0.3, lowess_it=2)
    t, r = port.synth_target(30000, 1), port.synth_reference(28000, 2)
    outs, st, fir, _, _ = run_pipeline(cfg, t, r)
    want = port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True)
    _compare(outs, want)


def test_pipeline_with_second_order_limiter_filters(lib):
    """The whole pipeline with LimiterConfig(hold_filter_order=2, release_filter_order=2): mgb_finalize sizes the
    look-back words and picks the limiter kernel from the plan's filter orders."""
    lim = port.OracleLimiterConfig(hold_filter_order=2, release_filter_order=2, release=30.0)
    cfg = port.OracleConfig(fft_size=1024, max_piece_size=0.3, limiter=lim)
    t, r = port.synth_target(30000, 1), port.synth_reference(28000, 2)
    outs, st, _, _, _ = run_pipeline(cfg, t, r)
    want = port.main(t.astype(np.float64), r.astype(np.float64), cfg, True, True, True)
    _compare(outs, want)
    assert st.limiter_engaged == 1


@pytest.mark.parametrize("rate_in,rate_out,n", [(48000, 44100, 20011), (22050, 44100, 9000), (96000, 44100, 30000),
                                                (44100, 48000, 7001), (8000, 44100, 700)])
def test_resampler_matches_oracle(lib, rate_in, rate_out, n):
    """mgb_resample (csrc/resample.cu) against the oracle's restatement of resampy's kaiser_best resampler:
    same table, same index arithmetic, float64 accumulation; down- and up-sampling, edges included."""
    import resample as oracle_resample
    from matchering_b200 import resample as product
    rng = np.random.default_rng(n)
    x = rng.uniform(-0.9, 0.9, (n, 2)).astype(np.float32)
    want = oracle_resample.resample(x.astype(np.float64), rate_in, rate_out)
    frames_out = int(lib.mgb_resample_frames(n, rate_in, rate_out))
    assert frames_out == want.shape[0]
    # the product's own table (built without the oracle) must be the oracle's table
    ratio = float(rate_out) / rate_in
    product._TABLES.clear()
    num_bits = 2 ** product.PRECISION
    m = num_bits * product.NUM_ZEROS
    from scipy.signal.windows import kaiser
    win = product.ROLLOFF * np.sinc(product.ROLLOFF * np.linspace(0, product.NUM_ZEROS, num=m + 1)) * kaiser(2 * m + 1, product.BETA)[m:]
    ref_win, ref_bits = oracle_resample.kaiser_best_filter()
    assert np.array_equal(win, ref_win) and num_bits == ref_bits
    if ratio < 1:
        win = ratio * win
    pairs = aligned_copy(np.stack([win, np.diff(win, append=win[-1])], axis=1))
    xin, out = aligned_copy(x), aligned((frames_out, 2), np.float32)
    _native.check(lib, lib.mgb_resample(ptr(xin), n, rate_in, ptr(out), frames_out, rate_out, ptr(pairs), len(win), num_bits, None))
    assert np.abs(out - want).max() < 2e-7
    assert lib.mgb_resample(ptr(xin), n, rate_in, ptr(out), frames_out + 1, rate_out, ptr(pairs), len(win), num_bits, None) == _native.MGB_ERR_INVALID


def test_second_order_release_filter_over_many_chunks(lib):
    """The default 3-second release as an order-2 section: its pole pair sits 3e-5 apart, and weights of whole chunks
    (companion-matrix powers up to C^147456) must be exact -- repeated squaring in float64 was off by 3e-3 there and
    cost 2e-5 in the output; the tables come from the closed form in extended precision (limiter.cu SectionPowers),
    including the weights of whole look-back windows (a running product of C^147456 cost 5e-6 at 96 kHz)."""
    for sr, n in ((44100, 115000), (96000, 330000)):  # 25 chunks; 72 chunks = three look-back windows of 32
        lim = port.OracleLimiterConfig(release_filter_order=2, hold_filter_order=2 if sr == 96000 else 1)
        cfg = port.OracleConfig(internal_sample_rate=sr, limiter=lim)
        x = port.synth_limiter_input(n, seed=12)
        x[n // 3:n // 2] *= 0.05
        want = port.limit(x.astype(np.float64), cfg)
        for inclusive in (1, 0):  # 0: aggregates only, every look-back walks all its predecessors (on the device a
            lib.mgb_set_option(b"lookback_inclusive", inclusive)  # chunk often finds only aggregates for a while)
            try:
                out, _ = _limit(lib, x, cfg)
            finally:
                lib.mgb_set_option(b"lookback_inclusive", 1)
            assert np.abs(out - want).max() < 3e-7


def _limiter_gains(lib, x, cfg):
    params = limiter_params(plan_mod.limiter_constants(cfg))
    n = len(x)
    ws_bytes = int(lib.mgb_limiter_workspace_bytes(C.byref(params), n))
    ws = aligned((ws_bytes,), np.uint8)
    xin, out = aligned_copy(x, np.float32), aligned((n, 2), np.float32)
    engaged = aligned((4,), np.int32)
    _native.check(lib, lib.mgb_test_limiter_gains(C.byref(params), ptr(xin), ptr(out), n, ptr(ws), ws_bytes, ptr(engaged), None))
    return out


def test_limiter_gain_envelopes_match_golden(lib, golden):
    """The limiter's two scanned envelopes themselves -- the attack gain (filtfilt, hyrax.py:43-53) and the release
    gain (hold / release low-passes, hyrax.py:56-75) -- against the UNMODIFIED reference's private helpers
    (tests/golden/limiter.npz), not just the samples they end up scaling."""
    g = golden("limiter.npz")
    gains = _limiter_gains(lib, g["x"], port.OracleConfig())
    assert np.abs(gains[:, 0] - g["gain_attack"]).max() < 2e-7
    assert np.abs(gains[:, 1] - g["gain_release"]).max() < 2e-7
    # and with second-order sections (golden from the oracle's trace)
    lim = port.OracleLimiterConfig(hold_filter_order=2, release_filter_order=2, release=30.0)
    cfg = port.OracleConfig(limiter=lim)
    trace = {}
    port.limit(g["x"].astype(np.float64), cfg, trace)
    gains2 = _limiter_gains(lib, g["x"], cfg)
    assert np.abs(gains2[:, 0] - trace["g_att"]).max() < 2e-7
    assert np.abs(gains2[:, 1] - np.maximum(trace["hold_out"], trace["rel_out"])).max() < 2e-7
