"""ORACLE (test infrastructure only) -- generate tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference, imported through oracle/ref_shims.py) on seeded synthetic inputs.

Run in the build container only:  python oracle/make_golden.py
The reference ships no golden vectors or tests of its own, so these files
are the pin: oracle/port.py and the CUDA path are both checked against them.  LOWESS inside the
reference run is oracle/lowess.py (statsmodels is not installable here; see that file).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import port  # noqa: E402
from ref_shims import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    warnings.simplefilter("ignore")
    import_reference()
    from matchering import Config, stages
    from matchering.limiter import limit
    from matchering.limiter import hyrax
    from matchering.stage_helpers import match_frequencies as mf
    from matchering.stage_helpers import match_levels as ml
    from matchering import dsp

    os.makedirs(OUT, exist_ok=True)

    # ---- G1: whole pipeline, 4 target pieces / 3 reference pieces, limiter engaged ------------
    cfg = Config(max_piece_size=0.25)
    target = port.synth_target(40000, seed=11)
    reference = port.synth_reference(33333, seed=12, kind="loud")
    limited, plain, normalized = stages.main(target.astype(np.float64), reference.astype(np.float64), cfg,
                                             need_default=True, need_no_limiter=True,
                                             need_no_limiter_normalized=True)
    # intermediates through the reference's own private helpers
    ref_n, coef = ml.normalize_reference(reference.astype(np.float64), cfg)
    t_an = ml.analyze_levels(target.astype(np.float64), "target", cfg)
    r_an = ml.analyze_levels(ref_n, "reference", cfg)
    c0 = r_an[4] / max(cfg.min_value, t_an[4])
    fir_mid = mf.get_fir(t_an[2] * c0, r_an[2], "mid", cfg)
    fir_side = mf.get_fir(t_an[3] * c0, r_an[3], "side", cfg)
    np.savez_compressed(
        os.path.join(OUT, "pipeline_small.npz"),
        target=target, reference=reference, max_piece_size_s=0.25,
        limited=limited, no_limiter=plain, normalized=normalized.astype(np.float32),
        final_amplitude_coefficient=coef, rms_coefficient=c0,
        target_match_rms=t_an[4], reference_match_rms=r_an[4],
        target_divisions=t_an[5], target_piece=t_an[6], reference_divisions=r_an[5], reference_piece=r_an[6],
        fir_mid=fir_mid, fir_side=fir_side)

    # ---- G2: quiet reference (normalised, final coefficient < 1), limiter early-out -----------
    cfg2 = Config(max_piece_size=0.3)
    target2 = port.synth_target(30000, seed=21, kind="white")
    reference2 = (0.05 * port.synth_reference(30011, seed=22, kind="quiet")).astype(np.float32)
    lim2, plain2, _ = stages.main(target2.astype(np.float64), reference2.astype(np.float64), cfg2, True, True, False)
    np.savez_compressed(os.path.join(OUT, "pipeline_quiet_reference.npz"), target=target2, reference=reference2,
                        max_piece_size_s=0.3, limited=lim2, no_limiter=plain2)

    # ---- G3: limiter alone, several chunks of the CUDA kernel, default and 96 kHz windows ------
    x = port.synth_limiter_input(30000, seed=31)
    y44 = limit(x.astype(np.float64), Config())
    y96 = limit(x.astype(np.float64), Config(internal_sample_rate=96000))
    g = dsp.flip(1.0 / dsp.rectify(x.astype(np.float64), Config().threshold))
    att, slided = getattr(hyrax, "__process_attack")(np.copy(g), Config())
    rel = getattr(hyrax, "__process_release")(np.copy(slided), Config())
    np.savez_compressed(os.path.join(OUT, "limiter.npz"), x=x, y_44100=y44, y_96000=y96,
                        gain_attack=att.astype(np.float32), gain_release=rel.astype(np.float32),
                        envelope=slided.astype(np.float32))
    # ---- G4: preview creator: the loudest 6-s window of the result, clipped target, fades ---------
    from matchering import Result, preview_creator
    cfg4 = Config(internal_sample_rate=2000, preview_size=6, preview_analysis_step=2)
    n4 = 30000
    target4 = (2.5 * port.synth_target(n4, 41)).astype(np.float32)
    result4 = (port.synth_reference(n4, 42) * (0.2 + np.abs(np.sin(np.linspace(0, 9, n4))))[:, None]).astype(np.float32)
    saved = {}
    original_save = preview_creator.save
    preview_creator.save = lambda file, arr, sr, subtype, name: saved.__setitem__(name, arr.copy())
    try:
        preview_creator.create_preview(target4.astype(np.float64), result4.astype(np.float64), cfg4,
                                       Result("t.wav", "PCM_16"), Result("r.wav", "PCM_16"))
    finally:
        preview_creator.save = original_save
    index4 = port.preview_pieces(target4, result4, cfg4)[0]
    np.savez_compressed(os.path.join(OUT, "preview.npz"), target=target4, result=result4, sample_rate=2000,
                        preview_size_s=6, preview_analysis_step_s=2, index=index4, every=7,
                        target_piece=saved["target preview"][::7], result_piece=saved["result preview"][::7],
                        target_piece_head=saved["target preview"][:300], result_piece_tail=saved["result preview"][-300:])
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
