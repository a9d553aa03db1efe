"""The reference's resampler-quality study in batched native calls (the caller of
``simulate_product_resampler``).

``python/tools/evaluate_resampler_quality.py`` renders every stimulus of a configuration with its own native call -- 16
pass-band tones, 3 stop-band tones + swept noise, 2 image tones, 2 impulses, the 44.1 -> 48 -> 44.1 kHz round trip, two
60 s sample-count streams: 28 calls per configuration, three configurations.  Signals of one length and direction share
everything but their samples (frame positions, phase table), so ``evaluate_configuration`` hands each such group to ONE
``simulate_product_resampler_batch`` call (11 calls; eight signals share every phase-table load on the GPU) and computes
the tool's measurements and gates from the outputs: the same numbers the tool reports (`_evaluate_configuration`,
:377-470), held against the published report in tests/test_resampler_eval.py.

Host logic only: stimuli and measurements are the tool's definitions (cited per function); the offline reference filter
(`scipy.signal.resample_poly`, :140-168) is the tool's own comparison reference and stays what it is.
"""
from __future__ import annotations

import math
from collections.abc import Callable
from typing import Any

import numpy as np
from scipy.signal import firwin, resample_poly

CHUNK_SIZE = 1024
PASSBAND_FREQUENCIES_HZ = (50.0, 100.0, 1_000.0, 5_000.0, 10_000.0, 15_000.0, 18_000.0, 20_000.0)  # :27-36
STOPBAND_FREQUENCIES_HZ = (22_500.0, 23_000.0, 23_500.0)
UPSAMPLE_IMAGE_TONES_HZ = (20_500.0, 21_000.0)
GATES = {  # :39-50
    "max_passband_absolute_error_db": 0.25, "max_passband_ripple_db": 0.25, "max_offline_reference_magnitude_delta_db": 0.25,
    "max_downsample_alias_db": -60.0, "max_upsample_image_db": -60.0, "max_impulse_location_error_samples": 1.0,
    "min_roundtrip_snr_db": 40.0, "max_long_stream_count_error_samples": 0, "max_p99_deadline_fraction": 0.25,
    "max_block_deadline_fraction": 0.50,
}


def _db_ratio(numerator: float, denominator: float) -> float:  # :68-71
    return -300.0 if numerator <= 0.0 else 20.0 * math.log10(numerator / max(denominator, 1e-15))


def _rms(values: np.ndarray) -> float:
    return float(np.sqrt(np.mean(np.square(values, dtype=np.float64))))


def _sine(sample_rate: int, frequency_hz: float, duration_seconds: float) -> np.ndarray:  # :112-115
    frames = int(round(sample_rate * duration_seconds))
    return 0.5 * np.sin(2.0 * np.pi * frequency_hz * (np.arange(frames, dtype=np.float64) / sample_rate))


def _steady(values: np.ndarray, sample_rate: int) -> np.ndarray:  # :118-122
    margin = min(int(round(0.25 * sample_rate)), max(0, values.size // 4))
    return values if margin == 0 else values[margin:-margin]


def _tone_amplitude(values: np.ndarray, sample_rate: int, frequency_hz: float) -> float:  # :125-137
    values = _steady(values, sample_rate)
    if values.size == 0:
        return 0.0
    window = np.hanning(values.size)
    phase = np.exp(-2j * np.pi * frequency_hz * np.arange(values.size, dtype=np.float64) / sample_rate)
    coherent_gain = float(np.sum(window)) / values.size
    return float(2.0 * np.abs(np.sum(values * window * phase)) / (values.size * max(coherent_gain, 1e-15)))


def _offline_reference(samples: np.ndarray, input_rate: int, output_rate: int) -> np.ndarray:  # :140-168
    common = math.gcd(input_rate, output_rate)
    up, down = output_rate // common, input_rate // common
    taps = firwin(2 * 64 * max(up, down) + 1, 1.0 / max(up, down), window=("kaiser", 14.0))
    return np.asarray(resample_poly(samples, up, down, window=np.asarray(taps, dtype=np.float64)), dtype=np.float64)


def _shaped_noise(sample_rate: int, low_hz: float, high_hz: float, duration_seconds: float, seed: int, pink: bool) -> np.ndarray:
    """`_band_limited_noise` (:214-230) / `_pink_noise` (:318-335): seeded spectrum, inverse FFT, 0.2 RMS."""
    frames = int(round(sample_rate * duration_seconds))
    frequencies = np.fft.rfftfreq(frames, d=1.0 / sample_rate)
    mask = (frequencies >= low_hz) & (frequencies <= high_hz)
    rng = np.random.default_rng(seed)
    spectrum = np.zeros(frequencies.size, dtype=np.complex128)
    draw = rng.standard_normal(mask.sum()) + 1j * rng.standard_normal(mask.sum())
    spectrum[mask] = draw / np.sqrt(frequencies[mask]) if pink else draw
    values = np.fft.irfft(spectrum, n=frames)
    return values * (0.2 / max(_rms(values), 1e-15))


def evaluate_configuration(identifier: str, sinc_len: int, window: str, *, native_default: bool = False,
                           duration_seconds: int = 60, simulate_batch: Callable | None = None) -> dict[str, Any]:
    """`_evaluate_configuration` (:377-470) with one native call per group of equal-length signals.  Returns the tool's
    `configuration`, `status`, `checks` and `measurements` (summary rows as the tracked report keeps them, wall-clock
    timing keys left out) plus `native_calls`."""
    if simulate_batch is None:
        from . import mic_eq_core
        simulate_batch = mic_eq_core.simulate_product_resampler_batch
    calls = 0

    def run(signals, input_rate: int, output_rate: int):
        nonlocal calls
        calls += 1
        args = (None, None) if native_default else (sinc_len, window)  # `_run` (:86-109): defaults for the product arm
        out, delay, expected = simulate_batch(np.stack(signals), input_rate, output_rate, CHUNK_SIZE, *args)
        out = np.asarray(out, dtype=np.float64)
        if out.shape[1] < expected:
            raise ValueError(f"resampler returned {out.shape[1]} frames; expected at least {expected}")
        return out[:, :expected], delay, expected

    passband = []
    for input_rate, output_rate in ((44_100, 48_000), (48_000, 44_100)):  # `_passband_case` (:171-211)
        sources = [_sine(input_rate, f, 1.5) for f in PASSBAND_FREQUENCIES_HZ]
        outputs, _, _ = run(sources, input_rate, output_rate)
        gains, deltas = [], []
        for source, output in zip(sources, outputs):
            input_rms = _rms(_steady(source, input_rate))
            gain = _db_ratio(_rms(_steady(output, output_rate)), input_rms)
            reference_gain = _db_ratio(_rms(_steady(_offline_reference(source, input_rate, output_rate), output_rate)), input_rms)
            gains.append(gain)
            deltas.append(abs(gain - reference_gain))
        passband.append({"input_rate": input_rate, "output_rate": output_rate, "max_absolute_error_db": max(abs(g) for g in gains),
                         "ripple_db": max(gains) - min(gains), "max_offline_reference_magnitude_delta_db": max(deltas)})

    tones = [_sine(48_000, f, 2.0) for f in STOPBAND_FREQUENCIES_HZ]  # `_downsample_alias_case` (:233-271)
    outputs, _, _ = run(tones, 48_000, 44_100)
    tone_db = [_db_ratio(_rms(_steady(o, 44_100)), _rms(_steady(s, 48_000))) for s, o in zip(tones, outputs)]
    noise = _shaped_noise(48_000, 22_500.0, 23_900.0, 4.0, 0xA11A5, pink=False)
    noise_out, _, _ = run([noise], 48_000, 44_100)
    noise_db = _db_ratio(_rms(_steady(noise_out[0], 44_100)), _rms(_steady(noise, 48_000)))
    downsample_alias = {"swept_noise": {"input_band_hz": [22_500.0, 23_900.0], "attenuation_db": noise_db},
                        "worst_alias_db": max(noise_db, *tone_db)}

    tones = [_sine(44_100, f, 2.0) for f in UPSAMPLE_IMAGE_TONES_HZ]  # `_upsample_image_case` (:274-297)
    outputs, _, _ = run(tones, 44_100, 48_000)
    upsample_image = {"worst_image_db": max(
        _db_ratio(_tone_amplitude(o, 48_000, 44_100.0 - f), _tone_amplitude(o, 48_000, f)) for f, o in zip(UPSAMPLE_IMAGE_TONES_HZ, outputs))}

    impulse, long_stream = [], []
    for input_rate, output_rate in ((44_100, 48_000), (48_000, 44_100)):
        source = np.zeros(input_rate, dtype=np.float64)  # `_impulse_case` (:300-315)
        source[input_rate // 2] = 1.0
        outputs, delay, _ = run([source], input_rate, output_rate)
        peak_index = int(np.argmax(np.abs(outputs[0])))
        expected_location = (input_rate // 2) * output_rate / input_rate
        impulse.append({"input_rate": input_rate, "output_rate": output_rate, "reported_output_delay_samples": delay,
                        "reported_output_delay_ms": delay * 1_000.0 / output_rate, "impulse_peak_index": peak_index,
                        "expected_impulse_location": expected_location, "location_error_samples": abs(peak_index - expected_location)})
    for input_rate, output_rate in ((44_100, 48_000), (48_000, 44_100)):  # `_long_stream_and_timing_case` (:365-374), counts only
        frames = input_rate * duration_seconds
        outputs, delay, _ = run([np.zeros(frames, dtype=np.float64)], input_rate, output_rate)
        expected = int(round(frames * output_rate / input_rate))
        long_stream.append({"input_rate": input_rate, "output_rate": output_rate, "duration_seconds": duration_seconds,
                            "input_frames": frames, "output_frames": int(outputs.shape[1]), "expected_output_frames": expected,
                            "count_error_samples": int(outputs.shape[1] - expected), "reported_output_delay_samples": delay})

    source = _shaped_noise(44_100, 50.0, 20_000.0, 8.0, 0xA0D10, pink=True)  # `_roundtrip_case` (:338-362)
    up, delay_up, _ = run([source], 44_100, 48_000)
    down, delay_down, _ = run([up[0]], 48_000, 44_100)
    length = min(source.size, down.shape[1])
    source_mid, roundtrip_mid = source[4_096:length - 4_096], down[0, 4_096:length - 4_096]
    error = roundtrip_mid - source_mid
    roundtrip = {"stimulus": "deterministic 50 Hz-20 kHz equal-energy-per-octave noise",
                 "roundtrip_snr_db": _db_ratio(_rms(source_mid), _rms(error)), "max_absolute_error": float(np.max(np.abs(error))),
                 "input_frames": int(source.size), "upsampled_frames": int(up.shape[1]), "roundtrip_frames": int(down.shape[1]),
                 "reported_up_delay_samples": delay_up, "reported_down_delay_samples": delay_down}

    checks = {  # :400-452, the gates that do not read a wall clock
        "passband_absolute_error": all(c["max_absolute_error_db"] <= GATES["max_passband_absolute_error_db"] for c in passband),
        "passband_ripple": all(c["ripple_db"] <= GATES["max_passband_ripple_db"] for c in passband),
        "offline_reference_magnitude": all(
            c["max_offline_reference_magnitude_delta_db"] <= GATES["max_offline_reference_magnitude_delta_db"] for c in passband),
        "downsample_alias": downsample_alias["worst_alias_db"] <= GATES["max_downsample_alias_db"],
        "upsample_image": upsample_image["worst_image_db"] <= GATES["max_upsample_image_db"],
        "impulse_location": all(c["location_error_samples"] <= GATES["max_impulse_location_error_samples"] for c in impulse),
        "delay_accounting": all(a["reported_output_delay_samples"] == b["reported_output_delay_samples"]
                                for a, b in zip(impulse, long_stream, strict=True)),
        "roundtrip": roundtrip["roundtrip_snr_db"] >= GATES["min_roundtrip_snr_db"],
        "long_stream_count": all(abs(c["count_error_samples"]) <= GATES["max_long_stream_count_error_samples"] for c in long_stream),
    }
    return {
        "configuration": {"identifier": identifier, "sinc_len": sinc_len, "window": window, "native_default": native_default,
                          "interpolation": "cubic", "oversampling_factor": 256},
        "status": "passed" if all(checks.values()) else "failed",  # the two realtime gates are the caller's (a wall clock)
        "checks": checks,
        "measurements": {"passband_and_offline_reference": passband, "downsample_alias": downsample_alias,
                         "upsample_image": upsample_image, "impulse": impulse, "roundtrip": roundtrip,
                         "long_stream_and_timing": long_stream},
        "native_calls": calls,
    }
