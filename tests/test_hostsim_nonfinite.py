"""Non-finite intermediates: at 8 / 11.025 / 22.05 kHz the de-esser's band edges reach Nyquist and its
dynamic EQ can blow up to inf / NaN mid-chain.  The reference's limiter queue treats a NaN as a barrier
(dsp/limiter.rs:216-237: every entry compared with NaN is popped, `front.max(|x|)` ignores a NaN operand); the
product's sliding-window maxima must reproduce that -- the stage bodies walked on the host are bit-identical to the
oracle (metrics and audio, NaNs included) on both kernel sets."""
import numpy as np
import pytest

from audio_forge_b200 import abi
from oracle import pyoracle
from tests import hostsim
from tests.cases import metric_mismatches
from tests.signals import speech_like
from tools.fuzz_parity import random_case


def test_low_rate_deesser_blowups_match_the_oracle_bit_for_bit():
    rng = np.random.default_rng(1)
    nonfinite = compared = 0
    for i in range(60):
        _, _, bands, overrides = random_case(rng)
        fs = float(rng.choice([8000.0, 11025.0, 22050.0, 192000.0]))
        overrides.update(deesser_enabled=True, input_stage="none", compressor_auto_makeup_enabled=False)
        x = speech_like(int(rng.integers(2000, 9000)), seed=int(rng.integers(1 << 30)), fs=fs, level=float(rng.uniform(0.05, 1.2)))
        cand = abi.AfCandidate()
        for b in range(abi.NUM_BANDS):
            cand.bands[b] = bands[b]
        cand.settings = abi.make_settings(**overrides)
        try:
            m0, a0, _ = pyoracle.chain_render(x, fs, bands, cand.settings, return_audio=True)
        except pyoracle.OracleError:  # a typed band beyond this rate's Nyquist: rejected on both sides
            with pytest.raises(hostsim.HostsimError):
                hostsim.chain_sweep([x], fs, (abi.AfCandidate * 1)(cand), [0], [0], split=1)
            continue
        compared += 1
        nonfinite += int(not np.all(np.isfinite(a0)))
        for split in (0, 1):
            m1, a1, _ = hostsim.chain_sweep([x], fs, (abi.AfCandidate * 1)(cand), [0], [0], split=split, want_audio=True)
            assert metric_mismatches(m0, m1[0]) == {}, (i, fs, split)
            assert np.array_equal(a0, a1[0], equal_nan=True), (i, fs, split)
    assert compared >= 20 and nonfinite >= 2  # the draw really contains renders whose limiter input went non-finite


def test_nan_in_the_limiter_window_follows_the_queue():
    """Direct statement of the rule on the oracle's limiter: a NaN hides the samples before it from later windows."""
    L = pyoracle.lib()
    lim = L.orc_limiter_new(-6.0, 50.0, 48000.0, 2.0)
    x = np.zeros(400, dtype=np.float32)
    x[100] = 0.9        # above the 0.501 ceiling: would hold the gain down for 96 more samples
    x[110] = np.nan     # ... but the NaN pops it from the queue
    x[150] = 0.2
    L.orc_limiter_process(lim, pyoracle.fptr(x), x.size)
    gr = L.orc_limiter_peak_gr_and_reset(lim)
    L.orc_limiter_free(lim)
    assert gr > 5.0  # limited while 0.9 was in the window
    cand = abi.AfCandidate()
    bands = abi.default_bands()
    for b in range(abi.NUM_BANDS):
        cand.bands[b] = bands[b]
    cand.settings = abi.make_settings(use_typed_bands=True, compressor_enabled=False, limiter_ceiling_db=-6.0,
                                      limiter_careful_output_enabled=False)
    # through the chain door NaN inputs are sanitised (python_api.rs:517-520): both sides agree trivially, bit for bit
    y = np.zeros(4000, dtype=np.float32)
    y[100], y[110], y[150] = 0.9, np.nan, 0.2
    m0, a0, _ = pyoracle.chain_render(y, 48000.0, cand.bands, cand.settings, return_audio=True)
    m1, a1, _ = hostsim.chain_sweep([y], 48000.0, (abi.AfCandidate * 1)(cand), [0], [0], split=1, want_audio=True)
    assert metric_mismatches(m0, m1[0]) == {} and np.array_equal(a0, a1[0])
