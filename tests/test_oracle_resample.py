"""oracle/resample.py (restatement of resampy's kaiser_best resampler; parity unpinned: resampy is not
installable here) against what a band-limited resampler must do."""
import numpy as np

import resample as rs


def _tone(freq, sr, n):
    return np.sin(2 * np.pi * freq * np.arange(n) / sr)[:, None] * np.array([[1.0, 0.5]])


def test_filter_table_shape_and_normalisation():
    win, bits = rs.kaiser_best_filter()
    assert win.shape == (32769,) and bits == 512
    assert abs(win[0] - rs.ROLLOFF) < 1e-15 and abs(win[-1]) < 1e-7
    # zero crossings of the sinc every 512/rolloff entries
    z = 512 / rs.ROLLOFF
    assert abs(win[int(round(z))]) < 2e-3 and abs(win[int(round(2 * z))]) < 2e-3


def test_output_length_rule():
    x = np.zeros((1001, 2))
    assert rs.resample(x, 48000, 44100).shape == (int(1001 * 44100 / 48000), 2)
    assert rs.resample(x, 22050, 44100).shape == (2002, 2)


def test_pass_band_tone_keeps_frequency_and_amplitude():
    for sr_in, sr_out in ((48000, 44100), (22050, 44100), (96000, 44100)):
        n = sr_in // 5
        y = rs.resample(_tone(1000.0, sr_in, n), sr_in, sr_out)
        want = _tone(1000.0, sr_out, y.shape[0])
        core = slice(400, y.shape[0] - 400)  # away from the edges, where the filter runs out of samples
        assert np.abs(y[core] - want[core]).max() < 1e-3


def test_content_above_the_new_nyquist_is_removed():
    sr_in, sr_out = 96000, 44100
    n = sr_in // 5
    y = rs.resample(_tone(30000.0, sr_in, n), sr_in, sr_out)  # 30 kHz > 22.05 kHz
    assert np.abs(y[400:-400]).max() < 5e-4
