"""The reference's own Python contract tests for this path, restated against the B200 backend.

Each test cites the reference test it restates (python/tests/...).  They go through the reference-facing
module ``audio_forge_b200.mic_eq_core`` (same function names, arguments and error messages as the PyO3 core)
and therefore need the GPU.  Tests that drive the live ``AudioProcessor`` object or Qt widgets are out of
scope.
"""
import math

import numpy as np
import pytest

from audio_forge_b200 import headroom, mic_eq_core
from audio_forge_b200.mic_eq_core import (eq_magnitude_response, eq_magnitude_response_v2, simulate_auto_eq_chain,
                                          simulate_eq_v2)

pytestmark = pytest.mark.gpu

EQ_FREQUENCIES = (80.0, 160.0, 320.0, 640.0, 1280.0, 2500.0, 5000.0, 8000.0, 12000.0, 16000.0)
DEFAULT_BANDS = [(f, 0.0, 1.41) for f in EQ_FREQUENCIES]


def _typed_default_bands():
    names = ["low_shelf"] + ["bell"] * 8 + ["high_shelf"]
    return [(names[i], f, 0.0, 1.41, 12, True) for i, f in enumerate(EQ_FREQUENCIES)]


# ---- python/tests/test_eq_native_response.py:25-54 --------------------------------------------------------------
def test_native_eq_response_validates_contract():
    with pytest.raises(ValueError, match="expected 10 EQ bands"):
        eq_magnitude_response([1000.0], DEFAULT_BANDS[:-1], 48_000.0)
    with pytest.raises(ValueError, match="sample_rate"):
        eq_magnitude_response([1000.0], DEFAULT_BANDS, 0.0)
    with pytest.raises(ValueError, match="Nyquist"):
        eq_magnitude_response([24_001.0], DEFAULT_BANDS, 48_000.0)


@pytest.mark.parametrize(("band_index", "frequency_hz", "gain_db", "q"),
                         [(1, 160.0, -12.0, 0.1), (4, 1000.0, 6.0, 2.0), (7, 8000.0, 12.0, 10.0)])
def test_native_peaking_response_reaches_configured_center_gain(band_index, frequency_hz, gain_db, q):
    bands = list(DEFAULT_BANDS)
    bands[band_index] = (frequency_hz, gain_db, q)
    response = eq_magnitude_response([frequency_hz], bands, 48_000.0)
    assert response[0] == pytest.approx(gain_db, abs=1.0e-8)


# ---- python/tests/test_eq_filter_types.py:58-197 -----------------------------------------------------------------
@pytest.mark.parametrize("filter_type", ["high_pass", "low_pass"])
@pytest.mark.parametrize("slope", [12, 24, 36, 48])
def test_typed_pass_response_is_minus_three_db_at_cutoff(filter_type, slope):
    bands = _typed_default_bands()
    bands[4] = (filter_type, 2000.0, 0.0, 1.0, slope, True)
    response = eq_magnitude_response_v2([2000.0], bands, 48_000.0)
    assert response[0] == pytest.approx(-20.0 * math.log10(math.sqrt(2.0)), abs=1e-8)


def test_typed_notch_response_ignores_gain_and_nulls_center():
    bands = _typed_default_bands()
    bands[4] = ("notch", 1000.0, 12.0, 8.0, 12, True)
    response = eq_magnitude_response_v2([100.0, 1000.0, 10_000.0], bands, 48_000.0)
    assert response[1] < -150.0
    assert abs(response[0]) < 0.1
    assert abs(response[2]) < 0.1


def test_disabled_typed_band_is_flat_and_finite():
    bands = _typed_default_bands()
    bands[4] = ("high_pass", 20_000.0, 12.0, 10.0, 48, False)
    response = eq_magnitude_response_v2(np.geomspace(20.0, 20_000.0, 100).tolist(), bands, 48_000.0)
    np.testing.assert_allclose(response, 0.0, rtol=0.0, atol=1e-12)


def test_typed_band_validation_messages():
    """test_eq_filter_types.py:33-55 checks these messages on the live object; the same validator guards this path."""
    bands = _typed_default_bands()
    bands[2] = ("tilt", 1000.0, 0.0, 1.0, 12, True)
    with pytest.raises(ValueError, match="band 2 has unsupported EQ filter type: tilt"):
        eq_magnitude_response_v2([1000.0], bands, 48_000.0)
    bands = _typed_default_bands()
    bands[3] = ("high_pass", 1000.0, 0.0, 1.0, 18, True)
    with pytest.raises(ValueError, match=r"Band 3: slope 18 dB/octave is unsupported; expected one of \[12, 24, 36, 48\]"):
        eq_magnitude_response_v2([1000.0], bands, 48_000.0)
    bands = _typed_default_bands()
    bands[5] = ("bell", 1000.0, 12.5, 1.0, 12, True)
    with pytest.raises(ValueError, match=r"Band 5: gain 12.5 dB out of range \[-12, 12\]"):
        eq_magnitude_response_v2([1000.0], bands, 48_000.0)


def test_native_typed_eq_simulator_preserves_default_audio_exactly():
    phase = np.arange(48_000, dtype=np.float32)
    audio = (0.2 * np.sin(phase * np.float32(2.0 * np.pi * 997.0 / 48_000.0))).astype(np.float32)
    result = simulate_eq_v2(audio, 48_000.0, _typed_default_bands(), return_output_audio=True)
    np.testing.assert_array_equal(np.asarray(result["output_audio"], dtype=np.float32), audio)
    assert result["algorithmic_latency_samples"] == 0
    assert result["non_finite_output"] is False
    assert result["max_response_db"] == pytest.approx(0.0, abs=1e-12)


def test_native_typed_eq_simulator_rejects_non_finite_audio():
    audio = np.asarray([0.0, np.nan], dtype=np.float32)
    with pytest.raises(ValueError, match="finite samples"):
        simulate_eq_v2(audio, 48_000.0, _typed_default_bands())


def test_native_typed_eq_simulator_handles_steep_pass_filter():
    rng = np.random.default_rng(0xA0D10)
    audio = rng.normal(0.0, 0.1, 48_000).astype(np.float32)
    bands = _typed_default_bands()
    bands[0] = ("high_pass", 80.0, 0.0, 1.41, 48, True)
    result = simulate_eq_v2(audio, 48_000.0, bands)
    assert result["sample_count"] == audio.size
    assert result["runtime_ms"] > 0.0
    assert result["algorithmic_latency_samples"] == 0
    assert result["non_finite_output"] is False
    assert math.isfinite(float(result["output_true_peak"]))


def test_typed_eq_full_chain_engages_limiter_and_respects_true_peak_ceiling():
    sample_rate = 48_000
    time = np.arange(sample_rate * 2, dtype=np.float64) / sample_rate
    audio = (0.5 * np.sin(2.0 * np.pi * 1000.0 * time)).astype(np.float32)
    typed = [("bell", 1000.0 + index * 10.0, 12.0, 10.0, 12, True) for index in range(10)]
    result = simulate_auto_eq_chain(audio, float(sample_rate), DEFAULT_BANDS, {
        "eq_bands_v2": typed, "deesser_enabled": False, "compressor_enabled": False, "limiter_enabled": True,
        "limiter_careful_output_enabled": True})
    assert result["non_finite_output"] is False
    assert result["limiter_gain_reduction_db"] > 1.0
    assert result["output_true_peak_db"] <= result["limiter_effective_ceiling_db"] + 0.05


# ---- python/tests/test_auto_eq.py:968-1023 (through the batched caller) ---------------------------------------------
def test_25_headroom_validation_reduces_boosts_when_peak_headroom_is_insufficient():
    sample_rate = 48_000
    t = np.arange(sample_rate, dtype=float) / sample_rate
    audio = (0.62 * np.sin(2.0 * np.pi * 5000.0 * t)).astype(np.float32)
    eq_settings = {"band_freqs": list(EQ_FREQUENCIES), "band_gains": [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 9.0, 0.0, 0.0, 0.0],
                   "band_qs": [1.41] * 10, "validation_gain_scale": 1.0, "validation_confidence": 0.95,
                   "analysis_confidence": 0.95}
    chain_settings = {"compressor": {"enabled": False}, "deesser": {"enabled": False},
                      "limiter": {"enabled": True, "ceiling_db": -0.5, "careful_output_enabled": True}}
    validated = headroom.apply_headroom_validation_batch(audio, sample_rate, [eq_settings], chain_settings)[0]
    assert validated["headroom_gain_scale"] < 1.0
    assert max(validated["band_gains"]) < 9.0
    assert validated["headroom_validation"]["safe"]
    assert validated["headroom_validation"]["after"]["pre_limiter_true_peak_headroom_db"] >= 1.0


def test_26_headroom_validation_preserves_safe_correction():
    sample_rate = 48_000
    t = np.arange(sample_rate, dtype=float) / sample_rate
    audio = (0.05 * np.sin(2.0 * np.pi * 180.0 * t) + 0.02 * np.sin(2.0 * np.pi * 1200.0 * t)).astype(np.float32)
    eq_settings = {"band_freqs": list(EQ_FREQUENCIES), "band_gains": [0.0, 0.0, 0.0, 1.5, 1.0, 0.5, 0.0, 0.0, 0.0, 0.0],
                   "band_qs": [1.41] * 10, "validation_gain_scale": 1.0, "validation_confidence": 0.90,
                   "analysis_confidence": 0.90}
    validated = headroom.apply_headroom_validation_batch(audio, sample_rate, [eq_settings])[0]
    assert validated["headroom_gain_scale"] == 1.0
    assert np.allclose(validated["band_gains"], eq_settings["band_gains"])
    assert validated["headroom_validation"]["safe"]


# ---- python/tests/test_voice_setup.py:481-513: result-dict shape the callers rely on ---------------------------------
def test_chain_result_has_the_reference_keys():
    audio = (0.1 * np.sin(2.0 * np.pi * 440.0 * np.arange(9600) / 48_000.0)).astype(np.float32)
    result = simulate_auto_eq_chain(audio, 48_000.0, DEFAULT_BANDS, None)
    expected = {
        "input_sample_peak_db", "input_rms_db", "output_sample_peak_db", "pre_limiter_true_peak_db", "output_true_peak_db",
        "output_rms_db", "limiter_effective_ceiling_db", "sample_headroom_db", "pre_limiter_true_peak_headroom_db",
        "true_peak_headroom_db", "limiter_gain_reduction_db", "true_peak_limiter_gain_reduction_db",
        "true_peak_limited_events", "compressor_gain_reduction_db", "deesser_gain_reduction_db",
        "compressor_gain_reduction_median_db", "compressor_gain_reduction_p95_db", "compressor_gain_reduction_active_ratio",
        "active_output_gain_db", "silence_output_gain_db", "silence_level_delta_db", "compressor_pumping_score_db",
        "non_finite_output", "candidate_runtime_ms", "deesser_gain_reduction_median_db", "deesser_gain_reduction_p95_db",
        "analysis_block_ms", "active_analysis_threshold_db", "active_analysis_block_count", "processed_samples"}
    assert set(result) == expected  # python_api.rs:649-713
    assert result["analysis_block_ms"] == 20.0 and result["processed_samples"] == 9600
    assert result["limiter_effective_ceiling_db"] == -1.5  # careful output (control.rs:904-910)
    assert mic_eq_core.list_input_devices() == []
