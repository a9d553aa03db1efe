"""The reference's own Rust unit tests for this path, restated against the CPU oracle (same inputs, same expected
values and tolerances).  Together with the golden vector (tests/test_oracle_golden.py) they pin the oracle: the
parity tests compare the CUDA path with an oracle that has passed the reference's known-answer tests.

Each test names the Rust test it restates (paths relative to rust-core/src).
"""
import ctypes as C

import numpy as np
import pytest

from audio_forge_b200 import abi
from oracle import pyoracle

FS = 48000.0
LOW_SHELF, HIGH_SHELF, PEAKING, NOTCH, HIGH_PASS, LOW_PASS, BYPASS = range(7)


@pytest.fixture(scope="module")
def L():
    return pyoracle.lib()


def _f32(values):
    return np.ascontiguousarray(values, dtype=np.float32)


# ---- dsp/biquad.rs -------------------------------------------------------------------------------------------------------
def test_notch_nulls_its_center_frequency(L):  # biquad.rs:481-487
    b = L.orc_biquad_new(NOTCH, 1000.0, 0.0, 4.0, FS)
    assert L.orc_biquad_response_db(b, 1000.0) < -150.0
    assert abs(L.orc_biquad_response_db(b, 100.0)) < 0.1
    assert abs(L.orc_biquad_response_db(b, 10000.0)) < 0.1
    L.orc_biquad_free(b)


def test_bypass_is_exactly_flat(L):  # biquad.rs:489-496
    b = L.orc_biquad_new(BYPASS, 1000.0, 12.0, 10.0, FS)
    x = _f32([-1.0, -0.25, 0.0, 0.25, 1.0])
    y = x.copy()
    L.orc_biquad_process(b, pyoracle.fptr(y), y.size)
    assert np.array_equal(x, y)
    assert L.orc_biquad_response_db(b, 1000.0) == 0.0
    L.orc_biquad_free(b)


def test_biquad_q_zero_guard(L):  # biquad.rs:498-503
    b = L.orc_biquad_new(PEAKING, 1000.0, 0.0, 0.0, FS)
    y = _f32([0.25])
    L.orc_biquad_process(b, pyoracle.fptr(y), 1)
    assert np.isfinite(y[0])
    L.orc_biquad_free(b)


def test_biquad_crossfade_promotes_pending_coefficients(L):  # biquad.rs:505-518: 72 samples = 1.5 ms at 48 kHz
    b = L.orc_biquad_new(PEAKING, 1000.0, 0.0, 1.0, FS)
    L.orc_biquad_set_gain_db(b, 12.0)
    assert L.orc_biquad_is_crossfading(b) == 1
    y = _f32(np.full(71, 0.2))
    L.orc_biquad_process(b, pyoracle.fptr(y), y.size)
    assert L.orc_biquad_is_crossfading(b) == 1
    y = _f32([0.2])
    L.orc_biquad_process(b, pyoracle.fptr(y), 1)
    assert L.orc_biquad_is_crossfading(b) == 0
    fresh = L.orc_biquad_new(PEAKING, 1000.0, 12.0, 1.0, FS)
    got, want = np.zeros(5), np.zeros(5)
    L.orc_biquad_coeffs(b, pyoracle.dptr(got))
    L.orc_biquad_coeffs(fresh, pyoracle.dptr(want))
    assert np.max(np.abs(got - want)) < 1e-12
    L.orc_biquad_free(b)
    L.orc_biquad_free(fresh)


def test_magnitude_response_matches_peaking_center_gain(L):  # biquad.rs:546-550
    b = L.orc_biquad_new(PEAKING, 1000.0, 6.0, 2.0, FS)
    assert abs(L.orc_biquad_response_db(b, 1000.0) - 6.0) < 1e-9
    L.orc_biquad_free(b)


def test_target_response_updates_only_after_the_crossfade(L):  # biquad.rs:560-567 (live response during the fade)
    b = L.orc_biquad_new(PEAKING, 1000.0, 0.0, 2.0, FS)
    L.orc_biquad_set_gain_db(b, 6.0)
    assert abs(L.orc_biquad_response_db(b, 1000.0)) < 1e-9
    L.orc_biquad_free(b)


# ---- dsp/eq.rs ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["high_pass", "low_pass"])
@pytest.mark.parametrize("slope", [12, 24, 36, 48])
def test_butterworth_pass_filters_are_minus_three_db_at_cutoff(kind, slope):  # eq.rs:698-715
    bands = abi.typed_bands([(kind, 1000.0, 0.0, 0.7, slope, True)] +
                            [("bell", f, 0.0, 1.41, 12, False) for f in abi.DEFAULT_FREQUENCIES[1:]])
    got = pyoracle.eq_response([1000.0], bands, FS, typed=True)[0]
    assert abs(got - (-3.010299956639812)) < 1e-8


def test_butterworth_section_q(L):  # eq.rs:203-207: 1 / (2 cos((2k + 1) pi / (4 n)))
    for n in (1, 2, 3, 4):
        for k in range(n):
            assert abs(L.orc_butterworth_q(k, n) - 1.0 / (2.0 * np.cos((2 * k + 1) * np.pi / (4 * n)))) < 1e-15


def test_typed_notch_ignores_gain_and_nulls_center():  # eq.rs:685-696
    bands = abi.typed_bands([("notch", 1000.0, 12.0, 4.0, 12, True)] +
                            [("bell", f, 0.0, 1.41, 12, False) for f in abi.DEFAULT_FREQUENCIES[1:]])
    assert pyoracle.eq_response([1000.0], bands, FS, typed=True)[0] < -150.0


# ---- dsp/compressor.rs ---------------------------------------------------------------------------------------------
def test_soft_knee_exact_boundaries(L):  # compressor.rs:860-872
    c = L.orc_comp_new(-20.0, 4.0, 10.0, 200.0, 0.0, 12.0, FS)
    assert abs(L.orc_comp_compute_gain_reduction(c, -20.0 - 6.0)) < 1e-12
    assert abs(L.orc_comp_compute_gain_reduction(c, -20.0 + 6.0) - 6.0 * 0.75) < 1e-12
    L.orc_comp_free(c)


def test_detector_blend_uses_linear_domain(L):  # compressor.rs:874-882
    d = L.orc_comp_blended_detector_db(-6.0, -18.0)
    assert -18.0 < d < -6.0
    assert abs(L.orc_comp_blended_detector_db(-12.0, -12.0) + 12.0) < 1e-9
    assert np.isfinite(L.orc_comp_blended_detector_db(-160.0, -160.0))


def test_time_constant_to_coeff(L):  # dsp/util.rs:5-9
    assert L.orc_time_constant_to_coeff(10.0, FS) == np.exp(-1.0 / (0.010 * FS))
    assert L.orc_time_constant_to_coeff(0.0, FS) == np.exp(-1.0 / (0.000001 * FS))  # floor of 0.001 ms


# ---- dsp/limiter.rs ----------------------------------------------------------------------------------------------------
def test_lookahead_scales_with_sample_rate(L):  # limiter.rs:312-325
    for fs, want in ((44100.0, 88), (48000.0, 96), (96000.0, 192), (192000.0, 384), (384000.0, 768)):
        lim = L.orc_limiter_new(-0.5, 50.0, fs, 2.0)
        assert L.orc_limiter_lookahead_samples(lim) == want
        L.orc_limiter_free(lim)


def test_limiter_no_reduction_below_ceiling(L):  # limiter.rs:327-340
    lim = L.orc_limiter_new(-0.5, 50.0, FS, 2.0)
    y = _f32(np.full(2048, 0.5))
    L.orc_limiter_process(lim, pyoracle.fptr(y), y.size)
    assert np.allclose(y[96:], 0.5, atol=1e-6)  # the output is the input delayed by the lookahead
    assert L.orc_limiter_peak_gr_and_reset(lim) == 0.0
    L.orc_limiter_free(lim)


def test_limiter_never_exceeds_ceiling(L):  # limiter.rs:374-400 family: hard clamp at the ceiling
    lim = L.orc_limiter_new(-1.5, 50.0, FS, 2.0)
    rng = np.random.default_rng(3)
    y = _f32(rng.uniform(-1.5, 1.5, 8192))
    L.orc_limiter_process(lim, pyoracle.fptr(y), y.size)
    assert np.max(np.abs(y)) <= 10.0 ** (-1.5 / 20.0) + 1e-7
    assert L.orc_limiter_peak_gr_and_reset(lim) > 0.0
    L.orc_limiter_free(lim)


# ---- dsp/true_peak.rs ------------------------------------------------------------------------------------------------
def test_constant_signal_matches_sample_peak(L):  # true_peak.rs:404-411
    d = L.orc_tpd_new()
    x = _f32(np.full(16, 0.5))
    assert abs(L.orc_tpd_process(d, pyoracle.fptr(x), 16) - 0.5) < 1e-6
    L.orc_tpd_free(d)


def test_bandlimited_oversampling_detects_intersample_overshoot(L):  # true_peak.rs:413-422
    d = L.orc_tpd_new()
    x = np.zeros(64, dtype=np.float32)
    x[1] = x[2] = 1.0
    assert L.orc_tpd_process(d, pyoracle.fptr(x), 64) > 1.0
    L.orc_tpd_free(d)


def test_true_peak_limiter_attenuates_intersample_overs(L):  # true_peak.rs:438-456
    t = L.orc_tpl_new(C.c_float(FS), C.c_float(0.0), C.c_float(60.0))
    L.orc_tpl_set_ceiling_linear(t, C.c_float(1.0))
    block = np.zeros(96, dtype=np.float32)
    block[1] = block[2] = 1.0
    stats = np.zeros(4, dtype=np.float32)
    L.orc_tpl_process(t, pyoracle.fptr(block), 96, pyoracle.fptr(stats))
    d = L.orc_tpd_new()
    out_peak = max(L.orc_tpd_process(d, pyoracle.fptr(block), 96), L.orc_tpd_process(d, pyoracle.fptr(np.zeros(48, dtype=np.float32)), 48))
    assert stats[0] == 1.0 and stats[1] > 1.0 and stats[3] > 0.0
    assert out_peak <= 1.0 + 1e-4
    L.orc_tpd_free(d)
    L.orc_tpl_free(t)


def test_true_peak_limiter_is_near_transparent_below_ceiling_after_delay(L):  # true_peak.rs:458-470
    t = L.orc_tpl_new(C.c_float(FS), C.c_float(-1.5), C.c_float(60.0))
    block = _f32(np.full(32, 0.25))
    stats = np.zeros(4, dtype=np.float32)
    L.orc_tpl_process(t, pyoracle.fptr(block), 32, pyoracle.fptr(stats))
    assert stats[0] == 0.0
    assert np.all(np.abs(block[20:] - 0.25) < 1e-6)  # delay = 20 samples (true_peak.rs:11)
    L.orc_tpl_free(t)


# ---- audio/processor/python_api.rs -------------------------------------------------------------------------------------
def test_pumping_score_is_zero_for_steady_gain_reduction(L):  # python_api.rs:767-771
    trace = _f32(np.full(250, 3.0))
    assert L.orc_pumping_score(pyoracle.fptr(trace), trace.size, C.c_float(50.0)) == 0.0


def test_pumping_score_focuses_on_fast_gain_modulation(L):  # python_api.rs:773-790
    i = np.arange(500, dtype=np.float32)
    fast = _f32(np.float32(3.0) + np.sin(np.float32(2.0 * np.pi) * np.float32(4.0) * i / np.float32(50.0)))
    slow = _f32(np.float32(3.0) + np.sin(np.float32(2.0 * np.pi) * np.float32(0.2) * i / np.float32(50.0)))
    assert (L.orc_pumping_score(pyoracle.fptr(fast), 500, C.c_float(50.0))
            > 2.0 * L.orc_pumping_score(pyoracle.fptr(slow), 500, C.c_float(50.0)))


def test_percentile_is_linear_interpolation_of_the_sorted_values(L):  # python_api.rs:58-72
    v = _f32([5.0, 1.0, 3.0, 2.0, 4.0])
    assert L.orc_percentile_f32(pyoracle.fptr(v), 5, C.c_float(0.5)) == 3.0
    assert L.orc_percentile_f32(pyoracle.fptr(v), 5, C.c_float(0.0)) == 1.0
    assert L.orc_percentile_f32(pyoracle.fptr(v), 5, C.c_float(1.0)) == 5.0
    assert abs(L.orc_percentile_f32(pyoracle.fptr(v), 5, C.c_float(0.9)) - 4.6) < 1e-6


# ---- audio/processor/tests.rs, routing.rs --------------------------------------------------------------------------------
def test_input_cleanup_off_is_the_fixed_prefilter(L):  # tests.rs:551-572: DC block + 80 Hz HP, no hum / rumble, HP at 80 Hz
    t = np.arange(4096, dtype=np.float32) / np.float32(FS)
    two_pi = np.float32(2.0 * np.pi)
    x = _f32(np.float32(0.15) * np.sin(two_pi * np.float32(60.0) * t) + np.float32(0.08) * np.sin(two_pi * np.float32(220.0) * t)
             + np.float32(0.04) * np.sin(two_pi * np.float32(1200.0) * t))
    info = np.zeros(4, dtype=np.float32)
    y = x.copy()
    L.orc_input_stage_process(1, FS, pyoracle.fptr(y), y.size, pyoracle.fptr(info))
    # the same filter built from its parts: x - x1 + 0.995 y1 (f32), then the 80 Hz / Q 0.707 high-pass biquad
    dc = np.zeros_like(x)
    x1 = y1 = np.float32(0.0)
    for i, v in enumerate(x):
        y1 = np.float32(v - x1 + np.float32(0.995) * y1)
        x1 = v
        dc[i] = y1
    hp = L.orc_biquad_new(HIGH_PASS, 80.0, 0.0, 0.707, FS)
    L.orc_biquad_process(hp, pyoracle.fptr(dc), dc.size)
    L.orc_biquad_free(hp)
    assert np.array_equal(y, dc)
    assert info[1] == 0.0 and info[2] == 0.0 and info[3] == 80.0


def test_fractional_hum_tracker_uses_power_and_phase_continuity(L):  # routing.rs:615-641
    n = 12000 * 3  # three 250 ms windows
    t = np.arange(n, dtype=np.float32) / np.float32(FS)
    two_pi = np.float32(2.0 * np.pi)
    x = _f32(np.float32(0.08) * np.sin(two_pi * np.float32(50.37) * t) + np.float32(0.025) * np.sin(two_pi * np.float32(100.74) * t))
    info = np.zeros(3, dtype=np.float32)
    L.orc_cleanup_analyze(1, C.c_float(FS), pyoracle.fptr(x), n, pyoracle.fptr(info))
    assert info[1] == 1.0 and info[2] > 0.0
    assert abs(float(info[0]) - 50.37) < 0.8
