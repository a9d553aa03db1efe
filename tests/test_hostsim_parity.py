"""CPU check of the product's stage-kernel bodies (tests/hostsim) against the oracle.

The hostsim harness compiles audio_forge_b200/csrc/afsim_render.h + afsim_plan.cpp for the host and
walks the same chunk x stage schedule as the CUDA launcher.  With the host libm on both sides the
result must be BIT-EXACT with the oracle (audio, per-block rows and every metric), for any chunk
size / ring depth / EQ slice width.  This pins the planner, the state parking, the ring addressing,
the van Herk limiter window and the finalize reduction without a GPU.
"""
import numpy as np
import pytest

from audio_forge_b200 import abi
from oracle import pyoracle
from tests import hostsim
from tests.cases import CASES, FS, candidate, candidate_array, metric_mismatches
from tests.signals import golden_chain_input, speech_like

X = golden_chain_input(blocks=60)  # 28 800 samples, the reference golden test's signal


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("schedule", [(1024, 2, 5, 0), (264, 3, 10, 0), (40000, 2, 5, 0), (1024, 2, 5, 7), (264, 3, 10, 7),
                                      (2048, 4, 5, 5), (512, 2, 5, 2), (1024, 2, 5, 15)])
def test_stage_bodies_bit_exact_with_oracle(name, schedule):
    bands, overrides = CASES[name]
    settings = abi.make_settings(**overrides)
    m0, a0, r0 = pyoracle.chain_render(X, FS, bands, settings, return_audio=True, return_rows=True)
    chunk, slots, eq_k, split = schedule
    cands = candidate_array([candidate(bands, **overrides)])
    m1, a1, r1 = hostsim.chain_sweep([X], FS, cands, [0], [0], chunk=chunk, slots=slots, eq_k=eq_k, split=split,
                                     want_audio=True, want_rows=True)
    assert np.array_equal(a0, a1[0])
    assert np.array_equal(r0, r1[:, :, 0].T)
    assert metric_mismatches(m0, m1[0]) == {}


def test_multi_stream_batch_and_ragged_tail():
    """Several candidates x passages in one batch; length not a multiple of chunk, block or FIR group."""
    n = 9 * 960 + 437
    passages = [speech_like(n, seed=s, level=0.7) for s in range(3)]
    bands, overrides = CASES["legacy_eq"]
    cand_list = [candidate(bands, **dict(overrides, compressor_threshold_db=thr, compressor_ratio=ratio))
                 for thr in (-30.0, -18.0) for ratio in (2.0, 6.0)]
    cands = candidate_array(cand_list)
    pp = np.array([p for c in range(len(cand_list)) for p in range(3)], dtype=np.uint32)
    pc = np.array([c for c in range(len(cand_list)) for p in range(3)], dtype=np.uint32)
    for split in (0, 7):
        got, audio, _ = hostsim.chain_sweep(passages, FS, cands, pp, pc, chunk=512, slots=2, split=split, want_audio=True)
        _check_pairs(passages, cand_list, pp, pc, got, audio)


@pytest.mark.parametrize("input_stage", ["none", "dc_hp80", "strong"])
def test_shared_input_stage_fan_out(input_stage):
    """A candidate sweep renders its input stage once per distinct passage and copies it to the streams (split bit 4):
    bit-identical to every stream rendering its own; chunk / block / passage length mutually ragged."""
    n = 4 * 960 + 437
    passages = [_hum_signal(n, seed=s) if input_stage == "strong" else speech_like(n, seed=s, level=0.7) for s in range(3)]
    bands, overrides = CASES["legacy_eq"]
    cand_list = [candidate(bands, **dict(overrides, input_stage=input_stage, compressor_threshold_db=thr, deesser_enabled=de))
                 for thr in (-30.0, -18.0) for de in (False, False)]
    cands = candidate_array(cand_list)
    pp = np.array([p for c in range(len(cand_list)) for p in range(3)], dtype=np.uint32)
    pc = np.array([c for c in range(len(cand_list)) for p in range(3)], dtype=np.uint32)
    for split, chunk in ((16, 1000), (16 | 7, 512)):
        got, audio, rows = hostsim.chain_sweep(passages, FS, cands, pp, pc, chunk=chunk, slots=2, split=split, want_audio=True,
                                               want_rows=True)
        _check_pairs(passages, cand_list, pp, pc, got, audio)
        for i in range(pp.size):  # the input-level rows (row.0) came through the fan-out as well
            _, _, r0 = pyoracle.chain_render(passages[pp[i]], FS, cand_list[pc[i]].bands, cand_list[pc[i]].settings, return_rows=True)
            assert np.array_equal(r0, rows[:, :, i].T), i


def test_shared_input_and_eq_prefix_fan_out():
    """A compressor grid over one EQ setting: input stage AND EQ run once per distinct (passage, EQ) pair (split bits
    4 + 5); legacy EQ (72-sample fade-in) and typed EQ, EQ-first orders only."""
    n = 3 * 960 + 411
    passages = [speech_like(n, seed=40 + s, level=0.7) for s in range(2)]
    for name, extra in (("legacy_eq", {}), ("typed_pass", {"deesser_enabled": False})):
        bands, overrides = CASES[name]
        cand_list = [candidate(bands, **dict(overrides, **extra, compressor_threshold_db=thr, compressor_ratio=ratio))
                     for thr in (-35.0, -20.0) for ratio in (2.0, 5.0)]
        cands = candidate_array(cand_list)
        pp = np.array([p for c in range(len(cand_list)) for p in range(2)], dtype=np.uint32)
        pc = np.array([c for c in range(len(cand_list)) for p in range(2)], dtype=np.uint32)
        for split, chunk in ((48, 1000), (48 | 7, 512), (112, 1000), (112 | 6, 520), (112 | 7, 1024), (112 | 15, 264)):  # 112: + shared compressor front
            got, audio, _ = hostsim.chain_sweep(passages, FS, cands, pp, pc, chunk=chunk, slots=2, split=split, want_audio=True)
            _check_pairs(passages, cand_list, pp, pc, got, audio)


def test_shared_deesser_front_fan_out():
    """De-esser first with one detector configuration: its detector biquads / envelopes / levels / confidence targets
    (R_a + M_b) run once per distinct passage (split bits 4 + 7); per-stream targets, rebuilds and filters read them."""
    n = 40 * 480 - 77
    passages = [golden_chain_input(blocks=40)[:n].copy(), speech_like(n, seed=71, level=0.8)]
    bands, overrides = CASES["golden_like"]  # de-esser auto mode, de-esser before the EQ
    cand_list = [candidate(bands, **dict(overrides, deesser_auto_amount=amount, deesser_max_reduction_db=red,
                                         compressor_threshold_db=-30.0 + 4 * i))
                 for i, (amount, red) in enumerate([(0.3, 6.0), (0.9, 12.0), (0.6, 3.0)])]
    cand_list.append(candidate(bands, **dict(overrides, deesser_auto_enabled=False, deesser_threshold_db=-45.0)))
    cands = candidate_array(cand_list)
    pp = np.array([p for c in range(len(cand_list)) for p in range(2)], dtype=np.uint32)
    pc = np.array([c for c in range(len(cand_list)) for p in range(2)], dtype=np.uint32)
    # 256: R_c1 cut into conf / baseline smoothers -> target map -> reduction smoothers + hysteresis
    for split, chunk in ((16 | 128, 1000), (16 | 128 | 7, 520), (16 | 128 | 8, 1024), (256, 1000), (256 | 7, 264),
                         (16 | 128 | 256, 520), (16 | 128 | 256 | 15, 1024)):
        got, audio, rows = hostsim.chain_sweep(passages, FS, cands, pp, pc, chunk=chunk, slots=2, split=split, want_audio=True,
                                               want_rows=True)
        _check_pairs(passages, cand_list, pp, pc, got, audio)
    de_max = max(got[i].deesser_gain_reduction_db for i in range(pp.size))
    assert de_max > 0.05  # the de-esser really worked on these passages


def _check_pairs(passages, cand_list, pp, pc, got, audio):
    for i in range(pp.size):
        m0, a0, _ = pyoracle.chain_render(passages[pp[i]], FS, cand_list[pc[i]].bands, cand_list[pc[i]].settings,
                                          return_audio=True)
        assert np.array_equal(a0, audio[i]), i
        assert metric_mismatches(m0, got[i]) == {}, i


@pytest.mark.parametrize("n", [1, 7, 71, 72, 73, 959, 960, 961])
def test_tiny_inputs(n):
    """Shorter than the crossfade / one analysis block / the limiter latency."""
    x = speech_like(2000, seed=5, level=0.9)[:n].copy()
    for name in ("legacy_eq", "golden_like"):
        bands, overrides = CASES[name]
        m0, a0, _ = pyoracle.chain_render(x, FS, bands, abi.make_settings(**overrides), return_audio=True)
        for split in (0, 7):
            m1, a1, _ = hostsim.chain_sweep([x], FS, candidate_array([candidate(bands, **overrides)]), [0], [0],
                                            split=split, want_audio=True)
            assert np.array_equal(a0, a1[0])
            assert metric_mismatches(m0, m1[0]) == {}


def test_non_finite_input_is_zeroed():
    x = speech_like(4000, seed=2)
    x[100] = np.nan
    x[2000] = np.inf
    x[2001] = -np.inf
    bands, overrides = CASES["default_legacy"]
    m0, a0, _ = pyoracle.chain_render(x, FS, bands, abi.make_settings(**overrides), return_audio=True)
    m1, a1, _ = hostsim.chain_sweep([x], FS, candidate_array([candidate(bands, **overrides)]), [0], [0], want_audio=True)
    assert np.array_equal(a0, a1[0])
    assert metric_mismatches(m0, m1[0]) == {}


def test_other_sample_rates():
    for fs in (44100.0, 96000.0, 16000.0):
        x = speech_like(int(fs * 0.4), seed=3, fs=fs, level=0.8)
        bands = abi.legacy_bands([(f if f < fs / 2 - 100 else fs / 2 - 500, g, q) for f, g, q in
                                  [(80, 3, 1), (160, -2, 1.2), (320, 1, 1.41), (640, -4, 2), (1280, 2, 0.7), (2500, 5, 1),
                                   (5000, -6, 3), (7000, 4, 1), (7400, 2, 1), (7600, -3, 0.8)]])
        overrides = dict(compressor_makeup_gain_db=8.0, deesser_enabled=fs > 30000.0)
        m0, a0, _ = pyoracle.chain_render(x, fs, bands, abi.make_settings(**overrides), return_audio=True)
        m1, a1, _ = hostsim.chain_sweep([x], fs, candidate_array([candidate(bands, **overrides)]), [0], [0],
                                        want_audio=True)
        assert np.array_equal(a0, a1[0]), fs
        assert metric_mismatches(m0, m1[0]) == {}, fs


def _hum_signal(n, hum_hz=50.37, level=0.1, seed=9):
    """Speech-like passage + mains hum and its second harmonic at -26 dBFS + a rumble burst."""
    # (defined below the tests that use it; looked up at call time)
    x = speech_like(n, seed=seed, level=0.5).astype(np.float64)
    t = np.arange(n) / FS
    x += level * np.sin(2 * np.pi * hum_hz * t) + 0.5 * level * np.sin(2 * np.pi * 2 * hum_hz * t + 0.3)
    burst = (t > 1.6) & (t < 1.9)
    x += burst * 0.6 * np.sin(2 * np.pi * 31.0 * t)
    return x.astype(np.float32)


@pytest.mark.parametrize("mode", ["gentle", "strong"])
@pytest.mark.parametrize("schedule", [(960, 2, 0), (2000, 3, 7)])
def test_adaptive_input_cleanup_bit_exact_with_oracle(mode, schedule):
    """Hum tracker + notches + rumble-adaptive high-pass (routing.rs:55-648) in front of the chain."""
    x = _hum_signal(3 * 48000)
    bands, overrides = CASES["legacy_eq"]
    overrides = dict(overrides, input_stage=mode)
    m0, a0, r0 = pyoracle.chain_render(x, FS, bands, abi.make_settings(**overrides), return_audio=True, return_rows=True)
    chunk, slots, split = schedule
    m1, a1, r1 = hostsim.chain_sweep([x], FS, candidate_array([candidate(bands, **overrides)]), [0], [0], chunk=chunk,
                                     slots=slots, split=split, want_audio=True, want_rows=True)
    assert np.array_equal(a0, a1[0])
    assert np.array_equal(r0, r1[:, :, 0].T)
    assert metric_mismatches(m0, m1[0]) == {}
    # the stage really engaged on this signal: the oracle's tracker locked onto the 50.37 Hz line
    info = np.zeros(4, dtype=np.float32)
    y = x.copy()
    pyoracle.lib().orc_input_stage_process(2 if mode == "gentle" else 3, FS, pyoracle.fptr(y), y.size, pyoracle.fptr(info))
    assert info[1] == 1.0 and abs(float(info[0]) - 50.37) < 0.8  # routing.rs:616-641 tolerance
