"""Compressor auto makeup (dsp/compressor.rs:598-653,700-722) and simulate_auto_makeup_control
(python_api.rs:118-276) on the serial makeup stage R7 (afsim_split.h).

The loudness meter restates the third-party `ebur128` crate 0.1.10 (not in the reference tree), so this sub-path
is pinned on the oracle's restatement only ("parity unpinned"); what these tests pin is that
the product's stage bodies and kernels reproduce that restatement: bit-exact on the CPU harness, within the
north_star tolerances on the GPU.
"""
import numpy as np
import pytest

from audio_forge_b200 import abi
from oracle import pyoracle
from tests import hostsim
from tests.cases import CASES, FS, audio_within_tolerance, candidate, candidate_array, metric_mismatches
from tests.signals import golden_chain_input, speech_like

X = golden_chain_input(blocks=300)  # 3 s: the 400 ms loudness window fills and turns over several times
MAKEUP = dict(compressor_auto_makeup_enabled=True, compressor_target_lufs=-16.0)


def _vad(blocks, seed=1):
    rng = np.random.default_rng(seed)
    return np.clip(0.5 + 0.5 * np.sin(np.arange(blocks) * 0.05) + rng.normal(0.0, 0.1, blocks), 0.0, 1.0)


# ---- CPU: the product's stage bodies (hostsim) against the oracle, bit-exact ---------------------------------------

@pytest.mark.parametrize("fs", [48000.0, 44100.0, 16000.0, 96000.0])
@pytest.mark.parametrize("with_vad", [False, True])
def test_control_sim_bodies_bit_exact_with_oracle(fs, with_vad):
    """480-sample control blocks at rates where a block spans 1 (48k), 4 (44.1k: gcd 120) or 3 (16k: gcd 160)
    loudness-window slots; the capture ends in a short block that is not a whole slot."""
    x = X[: int(fs * 1.3) + 123]
    blocks = (x.size + 479) // 480
    vad = _vad(blocks) if with_vad else None
    st = abi.make_makeup_settings(threshold_db=-30.0, makeup_gain_db=2.0)
    t0, a0 = pyoracle.auto_makeup_control(x, fs, vad, -60.0, 0.8, st, return_audio=True)
    for chunk, direct in ((960, False), (4000, True)):
        t1, a1 = hostsim.makeup_control(x, fs, vad, -60.0, 0.8, st, chunk=chunk, slots=3, direct=direct, want_audio=True)
        assert np.array_equal(t0, t1), (chunk, np.abs(t0 - t1).max(axis=1))
        assert np.array_equal(a0, a1)
    assert t0[0].max() - t0[0].min() > 1.0  # the controller really moved the makeup on this capture


@pytest.mark.parametrize("kwargs", [dict(adaptive_release=False, release_ms=90.0), dict(sidechain_highpass_enabled=False),
                                    dict(vad_reliability=0.2, target_lufs=-12.0), dict(ratio=8.0, attack_ms=1.0)])
def test_control_sim_settings_variants(kwargs):
    x = speech_like(60000, seed=4, level=0.4)
    vad = _vad((x.size + 479) // 480, seed=3)
    st = abi.make_makeup_settings(**kwargs)
    t0, a0 = pyoracle.auto_makeup_control(x, FS, vad, -48.0, 0.6, st, return_audio=True)
    t1, a1 = hostsim.makeup_control(x, FS, vad, -48.0, 0.6, st, want_audio=True)
    assert np.array_equal(t0, t1)
    assert np.array_equal(a0, a1)


@pytest.mark.parametrize("fs", [48000.0, 44100.0])
@pytest.mark.parametrize("schedule", [(1024, 2, 7), (3000, 3, 7), (1024, 2, 15), (1024, 2, 0)])
def test_chain_with_auto_makeup_bit_exact_with_oracle(fs, schedule):
    """simulate_auto_eq_chain with compressor_auto_makeup_enabled: 20 ms analysis blocks = one window slot each."""
    x = X[: 100000 + 77]
    bands, overrides = CASES["golden_like"]
    overrides = dict(overrides, **MAKEUP)
    m0, a0, r0 = pyoracle.chain_render(x, fs, bands, abi.make_settings(**overrides), return_audio=True, return_rows=True)
    chunk, slots, split = schedule
    m1, a1, r1 = hostsim.chain_sweep([x], fs, candidate_array([candidate(bands, **overrides)]), [0], [0], chunk=chunk,
                                     slots=slots, split=split, want_audio=True, want_rows=True)
    assert np.array_equal(a0, a1[0])
    assert np.array_equal(r0, r1[:, :, 0].T)
    assert metric_mismatches(m0, m1[0]) == {}
    # not the manual-makeup render: the controller changed the output level
    m2, _, _ = pyoracle.chain_render(x, fs, bands, abi.make_settings(**dict(overrides, compressor_auto_makeup_enabled=False)))
    assert abs(m2.output_rms_db - m0.output_rms_db) > 0.5


def test_auto_makeup_stays_off_without_a_meter_for_the_rate():
    """dsp/compressor.rs:318-323: no loudness meter for 22.05 kHz -> the switch is ignored (manual makeup)."""
    fs = 22050.0
    x = speech_like(20000, seed=6, fs=fs)
    bands = abi.legacy_bands([(f, g, q) for f, g, q in [(80, 3, 1), (160, -2, 1.2), (320, 1, 1.41), (640, -4, 2),
                                                        (1280, 2, 0.7), (2500, 5, 1), (5000, -6, 3), (7000, 4, 1),
                                                        (7400, 2, 1), (7600, -3, 0.8)]])
    on = dict(compressor_makeup_gain_db=4.0, **MAKEUP)
    m0, a0, _ = pyoracle.chain_render(x, fs, bands, abi.make_settings(**on), return_audio=True)
    m1, a1, _ = hostsim.chain_sweep([x], fs, candidate_array([candidate(bands, **on)]), [0], [0], want_audio=True)
    assert np.array_equal(a0, a1[0])
    off = dict(on, compressor_auto_makeup_enabled=False)
    m2, a2, _ = hostsim.chain_sweep([x], fs, candidate_array([candidate(bands, **off)]), [0], [0], want_audio=True)
    assert np.array_equal(a1[0], a2[0])


def test_control_sim_argument_errors_mirror_the_reference():
    """python_api.rs:136-166 messages, raised by the reference-facing module before any GPU work."""
    from audio_forge_b200 import mic_eq_core
    x = np.zeros(1440, dtype=np.float32)
    with pytest.raises(ValueError, match="sample_rate must be positive and finite"):
        mic_eq_core.simulate_auto_makeup_control(x, 0.0, [], -50.0, 1.0)
    with pytest.raises(ValueError, match="noise evidence must be finite and reliability must be between 0 and 1"):
        mic_eq_core.simulate_auto_makeup_control(x, FS, [], -50.0, 1.5)
    with pytest.raises(ValueError, match="VAD probabilities must be finite and between 0 and 1"):
        mic_eq_core.simulate_auto_makeup_control(x, FS, [0.0, 2.0, 1.0], -50.0, 1.0)


# ---- GPU: libafsim.so through the C ABI ------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def sim():
    from audio_forge_b200 import native
    s = native.Simulator(0)
    yield s
    s.close()


TRACE_TOL = (0.01, 1e-6, 1e-6, 0.01, 0.01, 0.01)  # dB traces within 0.01 dB; activity / reliability are f32 copies


@pytest.mark.gpu
@pytest.mark.parametrize("fs", [48000.0, 44100.0, 16000.0])
@pytest.mark.parametrize("with_vad", [False, True])
def test_gpu_control_sim_matches_oracle(sim, fs, with_vad):
    x = X[: int(fs * 2.2) + 123]
    vad = _vad((x.size + 479) // 480) if with_vad else None
    st = abi.make_makeup_settings(threshold_db=-30.0, makeup_gain_db=2.0)
    t0, a0 = pyoracle.auto_makeup_control(x, fs, vad, -60.0, 0.8, st, return_audio=True)
    t1, a1 = sim.auto_makeup_control(x, fs, vad, -60.0, 0.8, st, return_audio=True)
    for k, tol in enumerate(TRACE_TOL):
        assert np.max(np.abs(t0[k] - t1[k])) <= tol, (abi.MAKEUP_TRACES[k], np.max(np.abs(t0[k] - t1[k])))
    assert audio_within_tolerance(a0, a1) <= 0.0


@pytest.mark.gpu
def test_gpu_control_sweep_of_ragged_captures(sim):
    """One GPU pass over captures of different lengths / settings / evidence."""
    caps = [speech_like(30000 + 1111 * k, seed=30 + k, level=0.3 + 0.1 * k) for k in range(5)]
    vads = [None if k % 2 else _vad((c.size + 479) // 480, seed=k) for k, c in enumerate(caps)]
    sets = [abi.make_makeup_settings(threshold_db=-20.0 - 3 * k, adaptive_release=bool(k % 2)) for k in range(5)]
    floors, rels = [-55.0 + k for k in range(5)], [1.0, 0.5, 0.0, 0.9, 0.3]
    traces, outs = sim.auto_makeup_sweep(caps, FS, vads, floors, rels, sets, return_audio=True)
    for k in range(5):
        t0, a0 = pyoracle.auto_makeup_control(caps[k], FS, vads[k], floors[k], rels[k], sets[k], return_audio=True)
        for j, tol in enumerate(TRACE_TOL):
            assert np.max(np.abs(t0[j] - traces[k][j])) <= tol, (k, abi.MAKEUP_TRACES[j])
        assert audio_within_tolerance(a0, outs[k]) <= 0.0, k


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["fused", "split"])
def test_gpu_chain_with_auto_makeup_matches_oracle(sim, path, monkeypatch):
    """The compressor of an auto-makeup batch always runs split (R1 .. M6 R7); `path` selects the limiter kernels."""
    monkeypatch.setenv("AFSIM_SPLIT", "1" if path == "fused" else "2")
    x = X[: 100000 + 77]
    for name in ("golden_like", "legacy_eq"):
        bands, overrides = CASES[name]
        settings = abi.make_settings(**dict(overrides, **MAKEUP))
        m0, a0, _ = pyoracle.chain_render(x, FS, bands, settings, return_audio=True)
        m1, a1 = sim.chain_render(x, FS, bands, settings, return_audio=True)
        assert audio_within_tolerance(a0, a1) <= 0.0, name
        assert metric_mismatches(m0, m1, tol_db=0.01) == {}, name


@pytest.mark.gpu
def test_gpu_reference_python_door_for_the_control_hook(sim):
    """test_auto_makeup_real_speech_tools.py:82-98 restated against the B200 backend."""
    from audio_forge_b200 import mic_eq_core
    result = mic_eq_core.simulate_auto_makeup_control(np.zeros(1440, dtype=np.float32), 48000.0, [0.0, 0.5, 1.0], -50.0, 1.0)
    assert result["control_block_size"] == 480
    assert len(result["makeup_gain_db"]) == 3
    assert len(result["activity"]) == 3
    assert result["p99_block_runtime_ms"] >= 0.0
    with pytest.raises(ValueError, match="expected 3 VAD probabilities at the 10 ms control cadence, got 2"):
        mic_eq_core.simulate_auto_makeup_control(np.zeros(1440, dtype=np.float32), 48000.0, [0.0, 0.5], -50.0, 1.0)


@pytest.mark.gpu
def test_gpu_auto_makeup_sweep_with_shared_input(sim):
    """72 auto-makeup candidates x one passage: whole-block chunks, the shared input stage and the split compressor
    with R7 in one batch; a sample of the streams against the oracle."""
    x = X[: 60000 + 321]
    bands, overrides = CASES["legacy_eq"]
    cand_list = [candidate(bands, **dict(overrides, input_stage="dc_hp80", compressor_threshold_db=-40.0 + 0.4 * i,
                                         compressor_target_lufs=-22.0 + 0.1 * i, **{k: v for k, v in MAKEUP.items()
                                                                                    if k != "compressor_target_lufs"}))
                 for i in range(72)]
    got, _ = sim.chain_sweep([x], FS, candidate_array(cand_list))
    for i in (0, 17, 40, 71):
        m0, _, _ = pyoracle.chain_render(x, FS, cand_list[i].bands, cand_list[i].settings)
        assert metric_mismatches(m0, got[i], tol_db=0.01) == {}, i
