"""Time-parallel EQ render of one long passage (csrc/afsim_eqscan.h: segment-local runs + a scan of the 2x2
state-space recurrence) against the serial cascade of the oracle.  Every sample is still produced by the DF2T
recurrence; only the segment start states carry the scan's rounding, so the render stays far inside the north_star
tolerance (1e-5 relative or -100 dBFS) -- typically within one f32 ulp -- while the batched path stays bit-exact."""
import numpy as np
import pytest

from audio_forge_b200 import abi
from oracle import pyoracle
from tests import hostsim
from tests.cases import CASES, FS, audio_within_tolerance
from tests.signals import speech_like


@pytest.mark.parametrize("name", ["typed_pass", "typed_worst_40_sections"])
@pytest.mark.parametrize("log2_len", [3, 6, 10])
def test_scan_walk_matches_serial_cascade(name, log2_len):
    x = speech_like(70000 + 333, seed=3, level=0.7)
    bands, _ = CASES[name]
    _, a0 = pyoracle.eq_render(x, FS, bands, return_audio=True)
    a1 = hostsim.eq_scan(x, FS, bands, log2_len)
    assert audio_within_tolerance(a0, a1) <= 0.0
    assert np.max(np.abs(a0.astype(np.float64) - a1)) <= 2.5e-7  # one or two f32 ulps at full scale


@pytest.mark.parametrize("n", [1, 63, 64, 65, 4097])
def test_scan_walk_short_and_ragged(n):
    x = speech_like(5000, seed=8, level=0.9)[:n].copy()
    bands, _ = CASES["typed_pass"]
    _, a0 = pyoracle.eq_render(x, FS, bands, return_audio=True)
    a1 = hostsim.eq_scan(x, FS, bands, 6)
    assert audio_within_tolerance(a0, a1) <= 0.0


def test_scan_walk_flat_eq_is_passthrough():
    x = speech_like(9000, seed=9)
    assert np.array_equal(hostsim.eq_scan(x, FS, abi.default_bands(), 6), x)


@pytest.fixture(scope="module")
def sim():
    from audio_forge_b200 import native
    s = native.Simulator(0)
    yield s
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["typed_pass", "typed_worst_40_sections"])
@pytest.mark.parametrize("n", [65536, 480000 + 77])
def test_gpu_long_passage_eq_render(sim, name, n):
    """afsim_eq_render switches to the scan at 65536 samples: audio within tolerance, statistics within 1e-6."""
    x = speech_like(n, seed=5, level=0.8)
    bands, _ = CASES[name]
    st0, a0 = pyoracle.eq_render(x, FS, bands, return_audio=True)
    st1, a1 = sim.eq_render(x, FS, bands, return_audio=True)
    assert audio_within_tolerance(a0, a1) <= 0.0
    assert st1.input_sample_peak == st0.input_sample_peak and st1.input_true_peak == st0.input_true_peak
    assert st1.sample_count == st0.sample_count and st1.non_finite_output == st0.non_finite_output
    for key in ("output_sample_peak", "output_true_peak", "input_rms", "output_rms"):
        a, b = getattr(st0, key), getattr(st1, key)
        assert abs(a - b) <= 1e-6 * max(abs(a), 1e-6), key
    assert abs(st0.max_response_db - st1.max_response_db) < 1e-9


@pytest.mark.gpu
def test_gpu_scan_threshold_keeps_short_renders_bit_exact(sim, monkeypatch):
    x = speech_like(30000, seed=6)
    bands, _ = CASES["typed_pass"]
    _, a0 = pyoracle.eq_render(x, FS, bands, return_audio=True)
    _, a1 = sim.eq_render(x, FS, bands, return_audio=True)
    assert np.array_equal(a0, a1)
    monkeypatch.setenv("AFSIM_EQ_SCAN_MIN", "1")  # force the scan on the same passage
    _, a2 = sim.eq_render(x, FS, bands, return_audio=True)
    assert audio_within_tolerance(a0, a2) <= 0.0
