"""Randomised differential test: the kernels (CPU emulator, same .cu sources) against the oracle port over seeded random
Configs, lengths and levels -- fft sizes, sample rates up to 192 kHz, piece sizes down to one frame, 0..5 correction
steps, limiter timings and filter orders, LOWESS robustness iterations, mono targets, digital silence.  Configs the
package rejects (plan.UnsupportedConfig) and configs whose pieces are shorter than fft_size (where the reference's own
STFT changes its frame length) are skipped.  tools/fuzz_emul.py runs the same generator for as many cases as asked."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from fuzz_emul import random_case, run_case  # noqa: E402


@pytest.mark.parametrize("seed", range(10))
def test_random_config_against_oracle(seed):
    case = random_case(np.random.default_rng(1000 + seed))
    outcome, err, desc = run_case(case)
    if outcome == "skipped":
        pytest.skip(desc)
    assert outcome == "ok" and err < 1e-5, desc
