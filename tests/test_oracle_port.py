"""oracle/port.py pinned: against the golden vectors generated from the unmodified reference
(always), and against the reference itself where /root/reference exists."""
import numpy as np
import pytest

import port


def cfg_for(npz):
    return port.OracleConfig(max_piece_size=float(npz["max_piece_size_s"]))


def test_pipeline_small_matches_golden(golden):
    g = golden("pipeline_small.npz")
    cfg = cfg_for(g)
    tr = {}
    lim, plain, norm = port.main(g["target"].astype(np.float64), g["reference"].astype(np.float64), cfg,
                                 True, True, True, trace=tr)
    assert np.abs(lim - g["limited"]).max() < 1e-12
    assert np.abs(plain - g["no_limiter"]).max() < 1e-12
    assert np.abs(norm - g["normalized"]).max() < 2e-7  # stored as float32
    assert tr["target"]["divisions"] == int(g["target_divisions"]) and tr["target"]["piece"] == int(g["target_piece"])
    assert tr["reference"]["divisions"] == int(g["reference_divisions"])
    assert abs(tr["c0"] - float(g["rms_coefficient"])) < 1e-12
    assert abs(tr["final_coef"] - float(g["final_amplitude_coefficient"])) < 1e-15
    assert np.abs(tr["firs"]["mid"] - g["fir_mid"]).max() < 1e-13
    assert np.abs(tr["firs"]["side"] - g["fir_side"]).max() < 1e-13


def test_quiet_reference_matches_golden(golden):
    g = golden("pipeline_quiet_reference.npz")
    lim, plain, _ = port.main(g["target"].astype(np.float64), g["reference"].astype(np.float64), cfg_for(g),
                              True, True, False)
    assert np.abs(lim - g["limited"]).max() < 1e-12
    assert np.abs(plain - g["no_limiter"]).max() < 1e-12


def test_limiter_matches_golden(golden):
    g = golden("limiter.npz")
    x = g["x"].astype(np.float64)
    tr = {}
    assert np.abs(port.limit(x, port.OracleConfig(), trace=tr) - g["y_44100"]).max() < 1e-13
    assert np.abs(port.limit(x, port.OracleConfig(internal_sample_rate=96000)) - g["y_96000"]).max() < 1e-13
    assert np.abs(tr["a_env"] - g["envelope"]).max() < 1e-7
    assert np.abs(tr["g_att"] - g["gain_attack"]).max() < 1e-7
    assert np.abs(np.maximum(tr["hold_out"], tr["rel_out"]) - g["gain_release"]).max() < 1e-7
    assert abs(np.abs(g["y_44100"]).max() - port.OracleConfig().threshold) < 1e-9


def test_limiter_early_out_returns_input_object():
    x = 0.3 * port.synth_limiter_input(5000, 1).astype(np.float64)
    assert port.limit(x, port.OracleConfig()) is x


# ---- against the live reference (build container only) -------------------------------------------
def test_identities_against_reference_helpers(reference_package):
    from matchering import Config, dsp
    from matchering.limiter import hyrax
    from matchering.stage_helpers import match_frequencies as mf
    rng = np.random.default_rng(5)
    g = np.abs(rng.standard_normal(5000)) * (rng.uniform(size=5000) > 0.7)
    for attack in (44, 45, 96):
        want = getattr(hyrax, "__sliding_window_fast")(g, attack, "attack")
        reach = (attack + 1 if not attack & 1 else attack) - 1
        assert np.array_equal(port.centred_max(g, reach), want)
    for hold in (44, 45, 96, 3):
        want = getattr(hyrax, "__sliding_window_fast")(g, hold, "hold")
        assert np.array_equal(port.trailing_max(g, hold), want)
    cfg = Config()
    att, slided = getattr(hyrax, "__process_attack")(np.copy(g), cfg)
    k = port.limiter_coefficients(port.config_from(cfg))
    assert np.abs(port.one_pole_forward_backward(slided, k["c"]) - att).max() < 1e-15
    pieces = rng.standard_normal((3, 20000))
    want = getattr(mf, "__average_fft")(pieces, 44100, 4096)
    flat = pieces.reshape(-1)
    got = port.average_spectrum(flat, 20000, np.ones(3, dtype=bool), 4096)
    assert np.abs(got - want).max() < 1e-15
    x = rng.standard_normal((1000, 2))
    mid, side = dsp.lr_to_ms(x)
    m2, s2 = port.mid_side(x)
    assert np.array_equal(mid, m2) and np.array_equal(side, s2)


@pytest.mark.parametrize("sr,seconds", [(44100, 6.0), (96000, 2.5)])
def test_main_against_reference(reference_package, sr, seconds):
    from matchering import Config, stages
    n = int(sr * seconds)
    t = port.synth_target(n, 3).astype(np.float64)
    r = port.synth_reference(n - 777, 4).astype(np.float64)
    cfg = Config(internal_sample_rate=sr, max_piece_size=1.0)
    want = stages.main(t, r, cfg, True, True, True)
    got = port.main(t, r, cfg, True, True, True)
    for a, b in zip(got, want):
        assert np.abs(a - b).max() < 1e-12


@pytest.mark.parametrize("n,whole", [(30000, False), (9000, True), (12000 + 4000 * 3, False)])
def test_preview_pieces_against_reference(reference_package, monkeypatch, n, whole):
    """oracle/port.py::preview_pieces against the unmodified create_preview (its two `save` calls captured)."""
    from matchering import Config, Result, preview_creator
    cfg = Config(internal_sample_rate=2000, preview_size=6, preview_analysis_step=2)
    saved = {}
    monkeypatch.setattr(preview_creator, "save", lambda file, arr, sr, subtype, name: saved.__setitem__(name, arr.copy()))
    target = 2.5 * port.synth_target(n, 41).astype(np.float64)
    result = port.synth_reference(n, 42).astype(np.float64) * (0.2 + np.abs(np.sin(np.linspace(0, 9, n))))[:, None]
    keep_t, keep_r = target.copy(), result.copy()
    preview_creator.create_preview(target, result, cfg, Result("t.wav", "PCM_16"), Result("r.wav", "PCM_16"))
    index, t_piece, r_piece = port.preview_pieces(keep_t, keep_r, cfg)
    assert np.array_equal(saved["target preview"], t_piece) and np.array_equal(saved["result preview"], r_piece)
    assert (len(r_piece) == n) == whole
    if not whole:
        assert r_piece[0].tolist() == [0.0, 0.0] and r_piece[-1].tolist() == [0.0, 0.0]


def test_port_limiter_against_full_size_reference_golden(golden):
    """The first 40 s of BASELINE config 5's hour against the unmodified reference's decimated output
    (tests/golden/c5_limiter_hour.npz): the limiter is causal apart from its 88-sample look-ahead, so
    a prefix run equals the hour's prefix away from its own end."""
    g = golden("c5_limiter_hour.npz")
    n, every = int(g["frames"]), int(g["every"])
    m = 44100 * 40
    x = port.synth_limiter_input_prefix(n, m, int(g["seed"]))
    small = port.synth_limiter_input(5000, 3)
    assert np.array_equal(port.synth_limiter_input_prefix(5000, 1200, 3), small[:1200])
    y = port.limit(x.astype(np.float64), port.OracleConfig())
    keep = (m - 8192) // every
    assert np.abs(y[::every][:keep] - g["rows"][:keep]).max() < 1e-12
