"""GPU parity: libafsim.so (sm_100a kernels through the C ABI) against the CPU oracle.

Tolerances are the north_star's: rendered samples within 1e-5 relative or -100 dBFS absolute; gain
reduction / true peak / level metrics within 0.01 dB; counts and decisions exact.  The only
arithmetic that is not bit-identical to the oracle is the device libm (log10 / exp10 / sqrt are
IEEE or <= 2 ulp); coefficients, time constants and every IEEE +,-,*,/ match bit for bit.
"""
import numpy as np
import pytest

from audio_forge_b200 import abi
from oracle import pyoracle
from tests.cases import CASES, FS, audio_within_tolerance, candidate, candidate_array, metric_mismatches
from tests.signals import golden_chain_input, speech_like

pytestmark = pytest.mark.gpu

TOL_DB = 0.01


@pytest.fixture(scope="module")
def sim():
    from audio_forge_b200 import native
    s = native.Simulator(0)
    yield s
    s.close()


X = golden_chain_input(blocks=100)  # 48 000 samples = 47 chunks of 1024


@pytest.mark.parametrize("path", ["fused", "split", "tail"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_chain_render_matches_oracle(sim, name, path, monkeypatch):
    """All kernel paths: one-thread-per-stream fused stages (large sweeps), the R/M split (few streams) with the
    limiter / true-peak tail as five stage kernels, and with the fused SM-local tail kernel (the default)."""
    monkeypatch.setenv("AFSIM_SPLIT", "1" if path == "fused" else "2")
    monkeypatch.setenv("AFSIM_TAIL", "2" if path == "tail" else "1")
    bands, overrides = CASES[name]
    settings = abi.make_settings(**overrides)
    m0, a0, _ = pyoracle.chain_render(X, FS, bands, settings, return_audio=True)
    m1, a1 = sim.chain_render(X, FS, bands, settings, return_audio=True)
    assert audio_within_tolerance(a0, a1) <= 0.0
    assert metric_mismatches(m0, m1, tol_db=TOL_DB) == {}


def test_eq_only_paths_are_bit_exact(sim):
    """No transcendental on the signal path -> the EQ render must equal the oracle bit for bit."""
    for name in ("typed_pass", "typed_worst_40_sections"):
        bands, _ = CASES[name]
        st0, a0 = pyoracle.eq_render(X, FS, bands, return_audio=True)
        st1, a1 = sim.eq_render(X, FS, bands, return_audio=True)
        assert np.array_equal(a0, a1), name
        for key in ("input_sample_peak", "output_sample_peak", "input_true_peak", "output_true_peak", "input_rms",
                    "output_rms", "sample_count", "non_finite_output"):
            assert getattr(st0, key) == getattr(st1, key), (name, key)
        assert abs(st0.max_response_db - st1.max_response_db) < 1e-9


def test_default_eq_is_bit_exact_passthrough(sim):
    """test_eq_filter_types.py:123-142: default (flat) typed bands leave the samples untouched."""
    x = speech_like(48000, seed=11)
    st, out = sim.eq_render(x, FS, abi.default_bands(), return_audio=True)
    assert np.array_equal(out, x)
    assert st.output_sample_peak == st.input_sample_peak


@pytest.mark.parametrize("path", ["fused", "split"])
def test_sweep_mixed_structures_and_lengths(sim, path, monkeypatch):
    """One call, several batches: candidates with different stage sets x passages of different lengths."""
    monkeypatch.setenv("AFSIM_SPLIT", "1" if path == "fused" else "2")
    passages = [speech_like(20000 + 777 * k, seed=k, level=0.5 + 0.1 * k) for k in range(3)]
    names = ["default_legacy", "golden_like", "typed_pass", "no_limiter", "dc_hp", "long_lookahead"]
    cand_list = [candidate(CASES[n][0], **CASES[n][1]) for n in names]
    cands = candidate_array(cand_list)
    metrics, audio = sim.chain_sweep(passages, FS, cands, return_audio=True)
    for c in range(len(cand_list)):
        for p in range(len(passages)):
            i = c * len(passages) + p
            m0, a0, _ = pyoracle.chain_render(passages[p], FS, cand_list[c].bands, cand_list[c].settings, return_audio=True)
            assert audio_within_tolerance(a0, audio[i]) <= 0.0, (names[c], p)
            assert metric_mismatches(m0, metrics[i], tol_db=TOL_DB) == {}, (names[c], p)


@pytest.mark.parametrize("path", ["fused", "split"])
def test_sweep_many_streams_against_threaded_oracle(sim, path, monkeypatch):
    """A 24 x 4 compressor grid (ragged last warp: 96 streams + explicit pair lists)."""
    monkeypatch.setenv("AFSIM_SPLIT", "1" if path == "fused" else "2")
    passages = [speech_like(14400, seed=20 + k, level=0.8) for k in range(4)]
    bands, overrides = CASES["legacy_eq"]
    grid = [(thr, ratio, att) for thr in (-40.0, -30.0, -20.0, -12.0) for ratio in (1.5, 3.0, 6.0) for att in (3.0, 25.0)]
    cand_list = [candidate(bands, **dict(overrides, compressor_threshold_db=t, compressor_ratio=r, compressor_attack_ms=a,
                                         compressor_adaptive_release=(i % 2 == 0)))
                 for i, (t, r, a) in enumerate(grid)]
    cands = candidate_array(cand_list)
    pp = np.array([p for c in range(len(grid)) for p in range(4)][:-1], dtype=np.uint32)  # 95 pairs
    pc = np.array([c for c in range(len(grid)) for p in range(4)][:-1], dtype=np.uint32)
    got, _ = sim.chain_sweep(passages, FS, cands, pp, pc)
    want = pyoracle.chain_sweep(passages, FS, cands, pp, pc, n_threads=8)
    for i in range(pp.size):
        assert metric_mismatches(want[i], got[i], tol_db=TOL_DB) == {}, i


def test_resident_sweep_relaunch_is_deterministic(sim):
    bands, overrides = CASES["golden_like"]
    cands = candidate_array([candidate(bands, **overrides)])
    sweep = sim.prepare_sweep([X], FS, cands)
    sweep.launch()
    first = abi.metrics_to_dict(sweep.collect()[0])
    sweep.launch()
    second = abi.metrics_to_dict(sweep.collect()[0])
    assert sweep.kernel_count > 0 and sweep.render_ms() > 0.0
    sweep.release()
    first.pop("candidate_runtime_ms"), second.pop("candidate_runtime_ms")
    assert first == second


@pytest.mark.parametrize("path", ["fused", "split"])
@pytest.mark.parametrize("n", [0, 1, 71, 72, 73, 960, 961])
def test_tiny_and_empty_inputs(sim, n, path, monkeypatch):
    monkeypatch.setenv("AFSIM_SPLIT", "1" if path == "fused" else "2")
    x = speech_like(2000, seed=5, level=0.9)[:n].copy()
    bands, overrides = CASES["golden_like"]
    settings = abi.make_settings(**overrides)
    m0, a0, _ = pyoracle.chain_render(x, FS, bands, settings, return_audio=True)
    m1, a1 = sim.chain_render(x, FS, bands, settings, return_audio=True)
    assert audio_within_tolerance(a0, a1) <= 0.0
    assert metric_mismatches(m0, m1, tol_db=TOL_DB) == {}


def test_non_finite_input_is_zeroed(sim):
    x = speech_like(6000, seed=2)
    x[100], x[2000], x[2001] = np.nan, np.inf, -np.inf
    bands, overrides = CASES["default_legacy"]
    settings = abi.make_settings(**overrides)
    m0, a0, _ = pyoracle.chain_render(x, FS, bands, settings, return_audio=True)
    m1, a1 = sim.chain_render(x, FS, bands, settings, return_audio=True)
    assert audio_within_tolerance(a0, a1) <= 0.0
    assert metric_mismatches(m0, m1, tol_db=TOL_DB) == {}


def test_eq_response_matches_oracle_and_reference_pins(sim):
    freqs = np.geomspace(20.0, 20000.0, 100)
    for name, typed in (("typed_pass", True), ("typed_worst_40_sections", True), ("legacy_eq", False)):
        bands, _ = CASES[name]
        want = pyoracle.eq_response(freqs, bands, FS, typed)
        got = sim.eq_response(freqs, bands, FS, typed)[0]
        assert np.max(np.abs(want - got)) < 1e-8, name
    # Butterworth cascades are -3.0103 dB at the cutoff for every slope (dsp/eq.rs:699-715, +-1e-8)
    for slope in (12, 24, 36, 48):
        bands = abi.typed_bands([("high_pass", 1000.0, 0, 0.7, slope, True)] +
                                [("bell", f, 0, 1.41, 12, False) for f in abi.DEFAULT_FREQUENCIES[1:]])
        got = sim.eq_response([1000.0], bands, FS, True)[0, 0]
        assert abs(got - (-3.010299956639812)) < 1e-8, slope


def test_validation_errors_mirror_the_reference(sim):
    bad = abi.typed_bands([("bell", 10.0, 0, 1, 12, True)] + [("bell", f, 0, 1.41, 12, True) for f in abi.DEFAULT_FREQUENCIES[1:]])
    with pytest.raises(ValueError, match=r"Band 0: frequency 10 Hz out of range \[20, 23999\]"):
        sim.eq_response([100.0], bad, FS, True)
    with pytest.raises(ValueError, match="sample_rate must be positive and finite"):
        sim.chain_render(X[:100], float("nan"), abi.default_bands(), abi.make_settings())
    x = X[:100].copy()
    x[3] = np.nan
    with pytest.raises(ValueError, match="audio must contain only finite samples"):
        sim.eq_render(x, FS, abi.default_bands())


def test_reference_python_door(sim):
    """The reference-facing module: same names / dict shape as mic_eq_core (python_api.rs:649-713)."""
    from audio_forge_b200 import mic_eq_core
    x = (0.62 * np.sin(2 * np.pi * 5000.0 * np.arange(24000) / FS)).astype(np.float32)  # test_auto_eq.py:968-1001
    bands = [(f, 0.0, 1.41) for f in abi.DEFAULT_FREQUENCIES]
    bands[6] = (5000.0, 9.0, 1.41)
    result = mic_eq_core.simulate_auto_eq_chain(x, FS, bands, {"compressor_enabled": False, "return_output_audio": True})
    m0, a0, _ = pyoracle.chain_render(x, FS, abi.legacy_bands(bands), abi.make_settings(compressor_enabled=False),
                                      return_audio=True)
    want = abi.metrics_to_dict(m0)
    assert set(want) | {"output_audio"} == set(result)
    assert audio_within_tolerance(a0, np.asarray(result["output_audio"], dtype=np.float32)) <= 0.0
    # headroom.py:278-289 decision inputs
    for key in ("pre_limiter_true_peak_headroom_db", "limiter_gain_reduction_db", "true_peak_limiter_gain_reduction_db"):
        assert abs(want[key] - result[key]) <= TOL_DB
    assert result["limiter_gain_reduction_db"] > 1.0  # +9 dB at 5 kHz on a 0.62 sine engages the limiter


def _hum_signal(n, hum_hz=50.37, level=0.1, seed=9):
    """Speech-like passage + mains hum and its second harmonic + a rumble burst."""
    x = speech_like(n, seed=seed, level=0.5).astype(np.float64)
    t = np.arange(n) / FS
    x += level * np.sin(2 * np.pi * hum_hz * t) + 0.5 * level * np.sin(2 * np.pi * 2 * hum_hz * t + 0.3)
    burst = (t > 1.6) & (t < 1.9)
    x += burst * 0.6 * np.sin(2 * np.pi * 31.0 * t)
    return x.astype(np.float32)


@pytest.mark.parametrize("path", ["fused", "split"])
@pytest.mark.parametrize("mode", ["gentle", "strong"])
def test_adaptive_input_cleanup_matches_oracle(sim, mode, path, monkeypatch):
    """49-61 Hz hum tracker + hum / harmonic notches + rumble-adaptive high-pass (routing.rs:55-648) in front of
    the chain.  f32 throughout; only atan2f / logf at window ends and sinf / cosf at notch retunes use the device
    libm, so the render stays inside the sample tolerance."""
    monkeypatch.setenv("AFSIM_SPLIT", "1" if path == "fused" else "2")
    x = _hum_signal(3 * 48000)
    bands, overrides = CASES["golden_like"]
    settings = abi.make_settings(**dict(overrides, input_stage=mode))
    m0, a0, _ = pyoracle.chain_render(x, FS, bands, settings, return_audio=True)
    m1, a1 = sim.chain_render(x, FS, bands, settings, return_audio=True)
    assert audio_within_tolerance(a0, a1) <= 0.0
    assert metric_mismatches(m0, m1, tol_db=TOL_DB) == {}


@pytest.mark.parametrize("input_stage", ["none", "strong"])
def test_shared_input_stage_is_bit_identical(sim, input_stage, monkeypatch):
    """Candidate sweeps (>= 64 streams, >= 4 per passage) render the input stage once per distinct passage and fan
    it out: metrics and audio identical, bit for bit, to every stream rendering its own input stage."""
    passages = [_hum_signal(24000 + 333, seed=s) for s in range(2)]
    bands, overrides = CASES["legacy_eq"]
    cand_list = [candidate(bands, **dict(overrides, input_stage=input_stage, compressor_threshold_db=-40.0 + 0.7 * i,
                                         deesser_enabled=True)) for i in range(36)]
    cands = candidate_array(cand_list)
    monkeypatch.setenv("AFSIM_SHARED_INPUT", "2")  # off
    m_own, a_own = sim.chain_sweep(passages, FS, cands, return_audio=True)
    monkeypatch.setenv("AFSIM_SHARED_INPUT", "1")
    m_shared, a_shared = sim.chain_sweep(passages, FS, cands, return_audio=True)
    for i in range(len(cand_list) * 2):
        d0, d1 = abi.metrics_to_dict(m_own[i]), abi.metrics_to_dict(m_shared[i])
        d0.pop("candidate_runtime_ms"), d1.pop("candidate_runtime_ms")
        assert d0 == d1, i
        assert np.array_equal(a_own[i], a_shared[i]), i
    m0, a0, _ = pyoracle.chain_render(passages[1], FS, cand_list[5].bands, cand_list[5].settings, return_audio=True)
    assert audio_within_tolerance(a0, a_shared[5 * 2 + 1]) <= 0.0
    assert metric_mismatches(m0, m_shared[5 * 2 + 1], tol_db=TOL_DB) == {}


@pytest.mark.parametrize("path", ["fused", "split"])
@pytest.mark.parametrize("name", ["legacy_eq", "typed_worst_40_sections"])
def test_shared_eq_prefix_is_bit_identical(sim, name, path, monkeypatch):
    """A compressor grid over ONE EQ setting: the input stage and the EQ run once per distinct (passage, EQ) pair and
    are fanned out -- with the fused kernels also the compressor front (sidechain, detector weight, instantaneous
    peak); results identical, bit for bit, to every stream rendering its own prefix."""
    monkeypatch.setenv("AFSIM_SPLIT", "1" if path == "fused" else "2")
    passages = [speech_like(20000 + 123 * k, seed=60 + k, level=0.7) for k in range(2)]
    passages[1] = passages[1][: passages[0].size].copy()
    bands, overrides = CASES[name]
    cand_list = [candidate(bands, **dict(overrides, compressor_threshold_db=-45.0 + 0.5 * i, compressor_ratio=2.0 + 0.05 * i))
                 for i in range(40)]
    cands = candidate_array(cand_list)
    monkeypatch.setenv("AFSIM_SHARED_INPUT", "2")  # off
    m_own, a_own = sim.chain_sweep(passages, FS, cands, return_audio=True)
    monkeypatch.setenv("AFSIM_SHARED_INPUT", "1")
    m_shared, a_shared = sim.chain_sweep(passages, FS, cands, return_audio=True)
    for i in range(len(cand_list) * 2):
        d0, d1 = abi.metrics_to_dict(m_own[i]), abi.metrics_to_dict(m_shared[i])
        d0.pop("candidate_runtime_ms"), d1.pop("candidate_runtime_ms")
        assert d0 == d1, i
        assert np.array_equal(a_own[i], a_shared[i]), i
    m0, a0, _ = pyoracle.chain_render(passages[0], FS, cand_list[7].bands, cand_list[7].settings, return_audio=True)
    assert audio_within_tolerance(a0, a_shared[7 * 2]) <= 0.0
    assert metric_mismatches(m0, m_shared[7 * 2], tol_db=TOL_DB) == {}


@pytest.mark.parametrize("path", ["by_size", "split"])
def test_cut_compressor_grid_is_bit_identical(sim, path, monkeypatch):
    """A large compressor grid over few (passage, EQ) pairs is cut into pieces of AFSIM_SUBBATCH streams that run
    one after another on shared ring buffers: metrics and audio identical, bit for bit, to the uncut batch, on the
    first launch and on a relaunch of the resident sweep."""
    if path == "split":
        monkeypatch.setenv("AFSIM_SPLIT", "2")
    passages = [speech_like(9000 + 7, seed=80 + k, level=0.7) for k in range(4)]
    bands, overrides = CASES["legacy_eq"]
    cand_list = [candidate(bands, **dict(overrides, compressor_threshold_db=-45.0 + 0.4 * i, compressor_ratio=2.0 + 0.04 * i,
                                         compressor_adaptive_release=bool(i % 2))) for i in range(50)]
    cands = candidate_array(cand_list)
    monkeypatch.setenv("AFSIM_SUBBATCH", "100000")
    m_whole, a_whole = sim.chain_sweep(passages, FS, cands, return_audio=True)
    monkeypatch.setenv("AFSIM_SUBBATCH", "64")  # 200 streams -> pieces of 64 / 64 / 64 / 8 (the last one too small to share)
    sweep = sim.prepare_sweep(passages, FS, cands, want_audio=True)
    for _ in range(2):
        sweep.launch()
        m_cut = sweep.collect()
        a_cut = [sweep.collect_audio(i) for i in range(sweep.n_pairs)]
        for i in range(len(cand_list) * 4):
            d0, d1 = abi.metrics_to_dict(m_whole[i]), abi.metrics_to_dict(m_cut[i])
            d0.pop("candidate_runtime_ms"), d1.pop("candidate_runtime_ms")
            assert d0 == d1, i
            assert np.array_equal(a_whole[i], a_cut[i]), i
    sweep.release()
    m0, a0, _ = pyoracle.chain_render(passages[3], FS, cand_list[11].bands, cand_list[11].settings, return_audio=True)
    assert audio_within_tolerance(a0, a_cut[11 * 4 + 3]) <= 0.0
    assert metric_mismatches(m0, m_cut[11 * 4 + 3], tol_db=TOL_DB) == {}


@pytest.mark.parametrize("path", ["fused", "split"])
def test_deesser_target_stage_cut_is_bit_identical(sim, path, monkeypatch):
    """The de-esser's R_c1 as one serial kernel and cut three ways (R_c1a -> M_c1b -> R_c1c, AFSIM_DE_CUT): metrics
    and audio identical bit for bit, auto and manual mode, staged and direct kernel variants, own and shared
    detector front (72 streams over two passages share it)."""
    monkeypatch.setenv("AFSIM_SPLIT", "1" if path == "fused" else "2")
    n = 30 * 480 + 211
    passages = [golden_chain_input(blocks=31)[:n].copy(), speech_like(n, seed=91, level=0.8)]
    bands, overrides = CASES["golden_like"]  # de-esser auto mode, de-esser before the EQ
    cand_list = [candidate(bands, **dict(overrides, deesser_auto_amount=0.2 + 0.02 * i, deesser_max_reduction_db=3.0 + 0.3 * i))
                 for i in range(30)]
    cand_list += [candidate(bands, **dict(overrides, deesser_auto_enabled=False, deesser_threshold_db=-50.0 + i,
                                          deesser_max_reduction_db=4.0 + i)) for i in range(6)]
    cands = candidate_array(cand_list)
    monkeypatch.setenv("AFSIM_DE_CUT", "1")
    m_one, a_one = sim.chain_sweep(passages, FS, cands, return_audio=True)
    monkeypatch.setenv("AFSIM_DE_CUT", "2")
    m_cut, a_cut = sim.chain_sweep(passages, FS, cands, return_audio=True)
    for i in range(len(cand_list) * 2):
        d0, d1 = abi.metrics_to_dict(m_one[i]), abi.metrics_to_dict(m_cut[i])
        d0.pop("candidate_runtime_ms"), d1.pop("candidate_runtime_ms")
        assert d0 == d1, i
        assert np.array_equal(a_one[i], a_cut[i]), i
    assert max(m_cut[i].deesser_gain_reduction_db for i in range(len(cand_list) * 2)) > 0.05
    for c, p in ((3, 0), (32, 1)):
        m0, a0, _ = pyoracle.chain_render(passages[p], FS, cand_list[c].bands, cand_list[c].settings, return_audio=True)
        assert audio_within_tolerance(a0, a_cut[c * 2 + p]) <= 0.0
        assert metric_mismatches(m0, m_cut[c * 2 + p], tol_db=TOL_DB) == {}
