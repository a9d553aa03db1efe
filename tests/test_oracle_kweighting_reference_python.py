"""The oracle's restated `ebur128` momentary meter (oracle/afsim_oracle.hpp: K-weighting re-derived per sample rate from
the analogue prototype, one folded 4th-order section, 400 ms window, -0.691 + 10 log10(mean square)) against the K-weighting
the reference's OWN Python carries: `python/mic_eq/analysis/voice_setup.py:127-158` filters with the BS.1770 coefficient table
at 48 kHz (`shelf_b / shelf_a / highpass_b / highpass_a`, :133-136, restated below digit for digit -- data, not code) through
`scipy.signal.lfilter` and maps a window's mean square with the same `-0.691 + 10 log10(.)`.  Two independent routes to the
same standard: agreement to ~1e-4 dB says the oracle's filter design, its folding into one section and its window are the
BS.1770 ones (the table is rounded to 14 digits; the meter returns f32)."""
import ctypes as C

import numpy as np
import pytest
from scipy.signal import lfilter

from oracle import pyoracle
from tests.signals import speech_like

FS = 48000
SHELF_B = np.asarray([1.53512485958697, -2.69169618940638, 1.19839281085285])  # voice_setup.py:133
SHELF_A = np.asarray([1.0, -1.69065929318241, 0.73248077421585])               # :134
HIGHPASS_B = np.asarray([1.0, -2.0, 1.0])                                       # :135
HIGHPASS_A = np.asarray([1.0, -1.99004745483398, 0.99007225036621])             # :136
WINDOW = int(0.4 * FS)


def _reference_python_momentary(x: np.ndarray) -> float:
    weighted = lfilter(HIGHPASS_B, HIGHPASS_A, lfilter(SHELF_B, SHELF_A, x.astype(np.float64)))  # :137-140
    mean_square = float(np.mean(np.square(weighted[-WINDOW:])))
    return -0.691 + 10.0 * np.log10(mean_square + 1e-12)  # :155-156


def _oracle_momentary(x: np.ndarray) -> float:
    L = pyoracle.lib()
    meter = L.orc_meter_new(FS)
    buf = np.ascontiguousarray(x, dtype=np.float32)
    for start in range(0, buf.size, 480):  # fed in control blocks as the compressor does
        block = buf[start:start + 480]
        L.orc_meter_process(meter, block.ctypes.data_as(C.POINTER(C.c_float)), block.size)
    value = float(L.orc_meter_momentary(meter))
    L.orc_meter_free(meter)
    return value


def _signals():
    t = np.arange(FS, dtype=np.float64) / FS
    rng = np.random.default_rng(11)
    yield "1 kHz tone at -20 dBFS", (0.1 * np.sin(2.0 * np.pi * 1000.0 * t)).astype(np.float32)
    yield "100 Hz tone", (0.2 * np.sin(2.0 * np.pi * 100.0 * t)).astype(np.float32)
    yield "8 kHz tone", (0.05 * np.sin(2.0 * np.pi * 8000.0 * t)).astype(np.float32)
    yield "white noise", (0.05 * rng.standard_normal(FS)).astype(np.float32)
    yield "speech-like", speech_like(FS, seed=3, level=0.4)


@pytest.mark.parametrize("name,x", list(_signals()))
def test_momentary_loudness_equals_the_reference_python_k_weighting(name, x):
    want, got = _reference_python_momentary(x), _oracle_momentary(x)
    assert abs(got - want) < 2e-4, (name, want, got)


def test_one_kilohertz_tone_reads_what_bs1770_says():
    x = (0.1 * np.sin(2.0 * np.pi * 1000.0 * np.arange(FS) / FS)).astype(np.float32)
    # -20 dBFS peak sine: mean square -23.01 dB, K-weighting +0.69 dB at 1 kHz, -0.691 offset -> -23.0 LUFS
    assert abs(_oracle_momentary(x) - (-23.0)) < 0.02 and abs(_reference_python_momentary(x) - (-23.0)) < 0.02
