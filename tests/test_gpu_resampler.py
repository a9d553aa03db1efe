"""Product resampler simulator on the GPU (afsim_product_resampler behind the C ABI, the reference-facing
`mic_eq_core.simulate_product_resampler`) against the oracle (oracle/resampler_oracle.py, pinned on the reference's
published resampler-quality report) and against the published numbers themselves.

Tolerance: 1e-12 absolute at full scale 1.0 (a 128 / 256-tap f64 dot product summed in another order than numpy's:
~1e-16 x sqrt(taps) expected); frame counts, delays and positions exact."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from audio_forge_b200 import mic_eq_core, native
from oracle import resampler_oracle as R

pytestmark = pytest.mark.gpu
TOL = 1e-12
GOLDEN = json.loads((Path(__file__).parent / "golden" / "resampler_report.json").read_text())["configurations"]


def _signal(kind, n, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return 0.3 * rng.standard_normal(n)
    if kind == "impulse":
        x = np.zeros(n)
        x[n // 2] = 1.0
        return x
    t = np.arange(n, dtype=np.float64)
    return 0.5 * np.sin(2.0 * np.pi * 0.0137 * t) + 0.25 * np.sin(2.0 * np.pi * 0.31 * t + 0.4)


CASES = [
    (44100, 48000, 30000, 1024, None, None, "noise"),
    (48000, 44100, 30000, 1024, None, None, "noise"),
    (44100, 48000, 44100, 1024, 128, "blackman_harris_squared", "tones"),
    (48000, 44100, 20000, 1024, 256, "blackman_harris_squared", "noise"),
    (44100, 48000, 22050, 1024, None, None, "impulse"),
    (44100, 48000, 5000, 300, None, None, "noise"),
    (44100, 48000, 3000, 64, None, None, "tones"),
    (44100, 48000, 700, 1, None, None, "noise"),
    (16000, 48000, 8000, 1024, None, None, "noise"),
    (96000, 44100, 30000, 1000, None, None, "tones"),
    (44100, 48000, 1, 1024, None, None, "impulse"),
]


@pytest.mark.parametrize("case", CASES)
def test_render_equals_the_oracle(case):
    rate_in, rate_out, n_in, chunk, sinc_len, window, kind = case
    x = _signal(kind, n_in, seed=n_in)
    got, delay, expected, timings = mic_eq_core.simulate_product_resampler(x, rate_in, rate_out, chunk, sinc_len, window)
    want, o_delay, o_expected, _ = R.simulate_product_resampler(x, rate_in, rate_out, chunk, sinc_len, window)
    assert (delay, expected, len(got)) == (o_delay, o_expected, want.size)
    assert len(got) >= expected + delay and len(timings) > 0
    assert np.max(np.abs(np.asarray(got) - want)) <= TOL


def test_empty_input_flushes_to_the_delay():
    got, delay, expected, _ = mic_eq_core.simulate_product_resampler([], 44100, 48000)
    assert expected == 0 and delay == 69 and len(got) >= 69 and not np.any(got)


def test_batch_equals_single_calls_bit_for_bit():
    # 13 signals: one group of eight, one of four and a single one -- the three instances of the kernel
    signals = np.stack([_signal("noise", 12000, seed=s) for s in range(12)] + [_signal("tones", 12000)])
    out, delay, expected = mic_eq_core.simulate_product_resampler_batch(signals, 48000, 44100)
    assert out.shape[0] == 13 and delay == 58 and expected == 11025
    for s in range(signals.shape[0]):
        single, _, _, _ = mic_eq_core.simulate_product_resampler(signals[s], 48000, 44100)
        assert np.array_equal(out[s], np.asarray(single))


def test_batch_of_the_long_filter_equals_single_calls_and_the_oracle():
    signals = np.stack([_signal("noise", 9000, seed=40 + s) for s in range(9)])
    out, _, _ = mic_eq_core.simulate_product_resampler_batch(signals, 44100, 48000, 1024, 256, "blackman_harris_squared")
    for s in (0, 7, 8):
        single, _, _, _ = mic_eq_core.simulate_product_resampler(signals[s], 44100, 48000, 1024, 256, "blackman_harris_squared")
        assert np.array_equal(out[s], np.asarray(single))
    want, _, _, _ = R.simulate_product_resampler(signals[3], 44100, 48000, 1024, 256, "blackman_harris_squared")
    assert np.max(np.abs(out[3] - want)) <= TOL


def test_linearity_and_silence():
    a, b = _signal("noise", 20000, 1), _signal("tones", 20000)
    out, _, _ = mic_eq_core.simulate_product_resampler_batch(np.stack([a, b, a + b, np.zeros(20000)]), 44100, 48000)
    assert np.max(np.abs(out[0] + out[1] - out[2])) < 1e-13
    assert not out[3].any()


def test_device_resident_call_equals_the_host_call():
    import torch
    sim = mic_eq_core.simulator()
    signals = np.stack([_signal("noise", 50000, seed=s) for s in range(3)])
    spec = native.resampler_spec(44100, 48000)
    want, shape = sim.product_resampler(signals, spec)
    d_in = torch.from_numpy(signals).cuda()
    d_out = torch.zeros((3, int(shape.frames) + 5), dtype=torch.float64, device="cuda")
    ms = sim.product_resampler_device(spec, d_in.data_ptr(), signals.shape[1], 3, signals.shape[1], d_out.data_ptr(), d_out.shape[1])
    assert ms > 0.0
    assert np.array_equal(d_out[:, :int(shape.frames)].cpu().numpy(), want)
    assert not d_out[:, int(shape.frames):].any()


def test_reference_unit_test_invalid_inputs_through_the_door():
    # rust-core/src/audio/processor/tests.rs:222-259
    f = mic_eq_core.simulate_product_resampler
    with pytest.raises(ValueError, match="sample rates must be positive"):
        f([0.0], 0, 48000, 1024, None, None)
    with pytest.raises(ValueError, match="samples must be finite"):
        f([float("nan")], 48000, 44100, 1024, None, None)
    with pytest.raises(ValueError, match="chunk_size must be between 1 and 1024"):
        f([0.0], 48000, 44100, 0, None, None)
    with pytest.raises(ValueError, match="chunk_size must be between 1 and 1024"):
        f([0.0], 48000, 44100, 1025, None, None)
    with pytest.raises(ValueError, match="sinc_len must be a power of two between 32 and 2048"):
        f([0.0], 48000, 44100, 1024, 96, None)
    with pytest.raises(ValueError, match='unsupported resampler window "unknown"'):
        f([0.0], 48000, 44100, 1024, None, "unknown")
    with pytest.raises(RuntimeError, match="resampler flush produced no frames"):
        f(np.zeros(700), 48000, 44100, 1)
    with pytest.raises(native.AfsimError):
        f([0.0], 48000, 44100, 1024, 64, "hann")
    # tests.rs:193-206
    out, delay, expected, timings = f(np.zeros(44100), 44100, 48000, 1024, None, None)
    assert expected == 48000 and delay == 69 and len(out) >= delay + expected and timings


# ---- the reference's published numbers, rendered on the GPU (python/tools/evaluate_resampler_quality.py restated) ----
def _steady(v, sr):  # :118-122
    m = min(int(round(0.25 * sr)), max(0, v.size // 4))
    return v[m:-m] if m else v


def _rms(v):
    return float(np.sqrt(np.mean(np.square(v, dtype=np.float64))))


def _pink_noise(sample_rate, duration_seconds, low_hz, high_hz, seed):  # :318-335
    frames = int(round(sample_rate * duration_seconds))
    frequencies = np.fft.rfftfreq(frames, d=1.0 / sample_rate)
    mask = (frequencies >= low_hz) & (frequencies <= high_hz)
    rng = np.random.default_rng(seed)
    spectrum = np.zeros(frequencies.size, dtype=np.complex128)
    spectrum[mask] = (rng.standard_normal(mask.sum()) + 1j * rng.standard_normal(mask.sum())) / np.sqrt(frequencies[mask])
    values = np.fft.irfft(spectrum, n=frames)
    return values * (0.2 / max(_rms(values), 1e-15))


def test_published_stop_band_attenuation_is_reproduced_on_the_gpu():
    # _downsample_alias_case: the three stop-band tones in one batch; the worst one is the published worst_alias_db
    t = np.arange(96000, dtype=np.float64) / 48000
    tones = np.stack([0.5 * np.sin(2.0 * np.pi * f * t) for f in (22500.0, 23000.0, 23500.0)])
    out, _, expected = mic_eq_core.simulate_product_resampler_batch(tones, 48000, 44100)
    db = [20.0 * math.log10(_rms(_steady(out[i, :expected], 44100)) / _rms(_steady(tones[i], 48000))) for i in range(3)]
    published = GOLDEN["product"]["/measurements/downsample_alias/worst_alias_db"]["published"]
    assert abs(max(db) - published) < 1e-9, (db, published)


@pytest.mark.parametrize("name,sinc_len,window", [("product", None, None),
                                                  ("high-rejection-blackman-harris-squared-256", 256, "blackman_harris_squared")])
def test_published_round_trip_is_reproduced_on_the_gpu(name, sinc_len, window):
    # _roundtrip_case: 8 s of pink noise 44.1 -> 48 -> 44.1 kHz
    source = _pink_noise(44100, 8.0, 50.0, 20000.0, 0xA0D10)
    up, _, e_up = mic_eq_core.simulate_product_resampler_batch(source[None, :], 44100, 48000, 1024, sinc_len, window)
    down, _, e_down = mic_eq_core.simulate_product_resampler_batch(up[:, :e_up], 48000, 44100, 1024, sinc_len, window)
    roundtrip = down[0, :e_down]
    length = min(source.size, roundtrip.size)
    a, b = source[4096:length - 4096], roundtrip[4096:length - 4096]
    snr = 20.0 * math.log10(_rms(a) / _rms(b - a))
    rows = GOLDEN[name]
    assert abs(snr - rows["/measurements/roundtrip/roundtrip_snr_db"]["published"]) < 1e-9
    assert abs(float(np.max(np.abs(b - a))) - rows["/measurements/roundtrip/max_absolute_error"]["published"]) < 1e-12
    assert (e_up, e_down) == (384000, 352800)
