"""The fused tail kernel (csrc/afsim_tail.cu: limiter -> true-peak limiter -> detector + output statistics in one
SM-local, TMA-fed kernel) against the five split stage kernels it replaces: per sample the operations and their order
are the same, so metrics AND rendered audio must be bit-identical (AFSIM_TAIL=1 keeps the split kernels, 2 forces the tail), for every
lookahead the shared-memory x ring supports, ragged lengths, partial sub-tiles and odd chunk sizes."""
import numpy as np
import pytest

from audio_forge_b200 import abi
from oracle import pyoracle
from tests.cases import FS, LEGACY_EQ, audio_within_tolerance, candidate, candidate_array, metric_mismatches
from tests.signals import speech_like

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sim():
    from audio_forge_b200 import native
    s = native.Simulator(0)
    yield s
    s.close()


def _cands():
    items = []
    for la in (0.1, 0.5, 2.0, 5.0, 10.0):
        for ceil, makeup in ((-1.5, 6.0), (-6.0, 12.0)):
            items.append(candidate(abi.legacy_bands(LEGACY_EQ), limiter_lookahead_ms=la, limiter_ceiling_db=ceil,
                                   compressor_makeup_gain_db=makeup, limiter_careful_output_enabled=False))
    items.append(candidate(abi.default_bands(), use_typed_bands=True, compressor_enabled=False, limiter_ceiling_db=-3.0))
    return candidate_array(items)


@pytest.mark.parametrize("n,chunk", [(48000, "1024"), (30013, "1024"), (4099, "264"), (20, "1024"), (1031, "1024")])
def test_tail_is_bit_identical_to_the_split_kernels(sim, n, chunk, monkeypatch):
    monkeypatch.setenv("AFSIM_CHUNK", chunk)
    monkeypatch.setenv("AFSIM_SPLIT", "2")
    passages = [1.4 * speech_like(n, seed=21 + k) for k in range(3)]
    cands = _cands()
    monkeypatch.setenv("AFSIM_TAIL", "1")
    want, want_audio = sim.chain_sweep(passages, FS, cands, return_audio=True)
    monkeypatch.setenv("AFSIM_TAIL", "2")
    got, got_audio = sim.chain_sweep(passages, FS, cands, return_audio=True)
    for i in range(len(cands) * 3):
        assert metric_mismatches(want[i], got[i]) == {}, i
        assert np.array_equal(want_audio[i], got_audio[i]), i


def test_tail_matches_oracle_with_audio(sim, monkeypatch):
    monkeypatch.setenv("AFSIM_TAIL", "2")
    x = 1.4 * speech_like(48000, seed=5)
    cands = _cands()
    for i in (2, 5, 9, 10):
        want, ref_audio, _ = pyoracle.chain_render(x, FS, cands[i].bands, cands[i].settings, return_audio=True)
        got, audio = sim.chain_render(x, FS, cands[i].bands, cands[i].settings, return_audio=True)
        assert audio_within_tolerance(ref_audio, audio) <= 0.0
        assert metric_mismatches(want, got, tol_db=0.01) == {}


def test_tail_on_a_large_fused_batch(sim, monkeypatch):
    """AFSIM_TAIL=2: the tail kernel behind the one-thread-per-stream stage kernels of a > 16384-stream batch."""
    from audio_forge_b200 import workloads
    n, n_streams = 12000, 20480
    cands = workloads.true_peak_candidates(1)
    out = {}
    for mode in ("1", "2"):
        monkeypatch.setenv("AFSIM_TAIL", mode)
        sweep = sim.prepare_synthetic_sweep(1, n_streams, n, FS, cands)
        sweep.launch()
        out[mode] = sweep.collect()
        sweep.release()
    for i in range(0, n_streams, 97):
        assert metric_mismatches(out["1"][i], out["2"][i]) == {}, i
