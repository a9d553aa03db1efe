"""Parity at the BASELINE passage lengths: one stream of each benchmark shape rendered over its
FULL length on the GPU and by the CPU oracle, audio returned -- the <= 2-ulp device libm has to survive 0.5 - 2.9 M
samples of envelope / gain recurrences inside the north_star tolerances (samples 1e-5 relative or -100 dBFS, metrics
0.01 dB, counts exact), and the caller's decision from the metrics must be the same."""
import numpy as np
import pytest

from audio_forge_b200 import abi, workloads
from oracle import pyoracle
from tests.cases import FS, audio_within_tolerance, metric_mismatches

pytestmark = pytest.mark.gpu
TOL_DB = 0.01


@pytest.fixture(scope="module")
def sim():
    from audio_forge_b200 import native
    s = native.Simulator(0)
    yield s
    s.close()


def _one(cands, i):
    return (abi.AfCandidate * 1)(cands[i])


def _check(sim, x, cand):
    want, ref_audio, _ = pyoracle.chain_render(x, FS, cand.bands, cand.settings, return_audio=True)
    got, audio = sim.chain_render(x, FS, cand.bands, cand.settings, return_audio=True)
    assert audio_within_tolerance(ref_audio, audio) <= 0.0
    assert metric_mismatches(want, got, tol_db=TOL_DB) == {}
    w, g = abi.metrics_to_dict(want), abi.metrics_to_dict(got)
    assert workloads.is_headroom_safe(w) == workloads.is_headroom_safe(g)  # headroom.py:278-289
    return want, got


def test_c2_headroom_candidate_30s_16_sections(sim):
    """C2: a typed candidate with two 48 dB/oct pass bands (16 sections) over the 30 s passage (1.44 M samples)."""
    x = workloads.speech_like(int(30 * FS), seed=100, level=0.5)
    cands = workloads.headroom_candidates(64, seed=1234)
    idx = next(i for i in range(64) if cands[i].bands[0].filter_type == abi.FILTER_IDS["high_pass"])
    _check(sim, x, cands[idx])


def test_c2_sweep_slice_30s_matches_oracle_decisions(sim):
    """Eight C2 candidates (one full headroom ladder + 1) x 30 s in one sweep: metrics and the first-safe-scale walk."""
    x = workloads.speech_like(int(30 * FS), seed=100, level=0.5)
    cands = workloads.headroom_candidates(8, seed=1234)
    got, _ = sim.chain_sweep([x], FS, cands)
    pp, pc = np.zeros(8, dtype=np.uint32), np.arange(8, dtype=np.uint32)
    want = pyoracle.chain_sweep([x], FS, cands, pp, pc, n_threads=8)
    safe_w = [workloads.is_headroom_safe(abi.metrics_to_dict(want[i])) for i in range(8)]
    safe_g = [workloads.is_headroom_safe(abi.metrics_to_dict(got[i])) for i in range(8)]
    assert safe_w == safe_g
    for i in range(8):
        assert metric_mismatches(want[i], got[i], tol_db=TOL_DB) == {}, i


def test_c4_true_peak_stream_60s(sim):
    """C4: one hot-noise stream over 60 s (2.88 M samples) through limiter -> true-peak limiter -> detector."""
    n = int(60 * FS)
    x = workloads.synthetic_noise_host(7, n)
    cand = workloads.true_peak_candidates(1)[0]
    want, _ = _check(sim, x, cand)
    assert want.limiter_gain_reduction_db > 0.3


def test_c5_full_chain_stream_10s(sim):
    """C5: hum cleanup (Strong) -> auto de-esser -> typed EQ -> compressor -> limiter -> true peak over 10 s."""
    x = workloads.add_hum(workloads.speech_like(int(10 * FS), seed=100, level=0.6), 50.37)
    cands = workloads.full_chain_candidates(16, seed=1234)
    for i in (0, 9):  # adaptive release off / on
        _check(sim, x, cands[i])


def test_c3_compressor_grid_stream_20s(sim):
    """C3: two grid corners over the 20 s passage, adaptive release, search limiter."""
    x = workloads.speech_like(int(20 * FS), seed=101, level=0.6)
    cands = workloads.compressor_grid_candidates(256, seed=1234)
    for i in (0, 255):
        _check(sim, x, cands[i])
