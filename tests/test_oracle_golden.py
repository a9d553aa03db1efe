"""Pins the oracle against the reference's golden known-answer test.

Restates rust-core/src/audio/processor/tests.rs:1784-1885
(test_full_downstream_chain_matches_golden_tolerance): same setters in the same
order on OfflineDspBlockProcessor, same LCG+sines input, 300 blocks of 480, same
expected values and tolerances.  Also tests.rs:1741-1781 (offline == manual stage
composition).
"""
import ctypes as C

import numpy as np

from oracle import pyoracle
from tests.signals import golden_chain_input

FS = 48000.0


def _run_golden():
    L = pyoracle.lib()
    p = L.orc_proc_new(FS)
    L.orc_proc_set_deesser_enabled(p, 1)
    L.orc_proc_set_eq_enabled(p, 1)
    L.orc_proc_set_compressor_enabled(p, 1)
    L.orc_proc_set_limiter_enabled(p, 1)
    L.orc_proc_deesser_set_auto_enabled(p, 1)
    L.orc_proc_deesser_set_auto_amount(p, 0.85)
    L.orc_proc_deesser_set_max_reduction_db(p, 10.0)
    for band, f, g, q in ((2, 180.0, -2.5, 0.8), (6, 2800.0, 3.0, 1.2), (8, 7200.0, 1.5, 1.0)):
        L.orc_proc_eq_set_band_frequency(p, band, f)
        L.orc_proc_eq_set_band_gain(p, band, g)
        L.orc_proc_eq_set_band_q(p, band, q)
    L.orc_proc_comp_set_threshold(p, -22.0)
    L.orc_proc_comp_set_ratio(p, 3.5)
    L.orc_proc_comp_set_attack_time(p, 8.0)
    L.orc_proc_comp_set_release_time(p, 160.0)
    L.orc_proc_comp_set_makeup_gain(p, 8.0)
    L.orc_proc_comp_set_adaptive_release(p, 1)
    L.orc_proc_limiter_set_ceiling(p, -6.0)
    L.orc_proc_limiter_set_release_time(p, 55.0)

    x = golden_chain_input()
    out = np.empty_like(x)
    stats = np.zeros(9, dtype=np.float32)
    agg = dict(comp=0.0, de=0.0, lim=0.0, events=0)
    for b in range(300):
        block = x[b * 480:(b + 1) * 480].copy()
        L.orc_proc_process_block(p, pyoracle.fptr(block), 480, pyoracle.fptr(stats))
        out[b * 480:(b + 1) * 480] = block
        agg["comp"] = max(agg["comp"], float(stats[6]))
        agg["de"] = max(agg["de"], float(stats[7]))
        agg["lim"] = max(agg["lim"], float(stats[4]), float(stats[5]))
        agg["events"] += int(stats[8])
    L.orc_proc_free(p)
    return out, agg


def test_full_downstream_chain_matches_reference_golden_vector():
    out, agg = _run_golden()
    o64 = out.astype(np.float64)
    rms = float(np.sqrt(np.sum(o64 * o64) / out.size))
    peak = float(np.max(np.abs(out)))
    weights = (np.arange(out.size) % 997 + 1).astype(np.float64)
    weighted = float(np.sum(o64 * weights))
    assert abs(rms - 0.185_715_270_552) <= 1.0e-6
    assert abs(peak - 0.500_814_14) <= 2.0e-6
    assert abs(weighted - (-4_246.481_547_342)) <= 0.05
    assert abs(agg["comp"] - 8.687_991) <= 0.001
    assert abs(agg["de"] - 10.0) <= 0.001
    assert abs(agg["lim"] - 4.348_602) <= 0.001
    assert 20 <= agg["events"] <= 24
    expected = [-0.038_492_45, 0.185_469_2, 0.200_082_9, -0.093_881_376]
    for index, value in zip((1_000, 10_000, 50_000, 100_000), expected):
        assert abs(float(out[index]) - value) <= 2.0e-5


def _rms_error_db(a, b):
    err = np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 20.0 * np.log10(max(err, 1e-12))


def test_offline_block_processor_matches_manual_stage_composition():
    # processor/tests.rs:1741-1781
    L = pyoracle.lib()
    idx = np.arange(512)
    t = idx / FS
    x = (0.38 * np.sin(2 * np.pi * 2500.0 * t) + 0.22 * np.sin(2 * np.pi * 180.0 * t)).astype(np.float32)

    p = L.orc_proc_new(FS)
    L.orc_proc_set_deesser_enabled(p, 0)
    L.orc_proc_set_eq_enabled(p, 1)
    L.orc_proc_set_compressor_enabled(p, 0)
    L.orc_proc_set_limiter_enabled(p, 1)
    L.orc_proc_eq_set_band_frequency(p, 5, 2500.0)
    L.orc_proc_eq_set_band_gain(p, 5, 4.0)
    L.orc_proc_eq_set_band_q(p, 5, 1.8)
    L.orc_proc_limiter_set_ceiling(p, -1.5)
    offline = x.copy()
    stats = np.zeros(9, dtype=np.float32)
    L.orc_proc_process_block(p, pyoracle.fptr(offline), 512, pyoracle.fptr(stats))
    L.orc_proc_free(p)

    # manual: EQ-only processor, then Limiter, then TruePeakLimiter objects
    q = L.orc_proc_new(FS)
    L.orc_proc_set_limiter_enabled(q, 0)
    L.orc_proc_eq_set_band_frequency(q, 5, 2500.0)
    L.orc_proc_eq_set_band_gain(q, 5, 4.0)
    L.orc_proc_eq_set_band_q(q, 5, 1.8)
    manual = x.copy()
    s2 = np.zeros(9, dtype=np.float32)
    L.orc_proc_process_block(q, pyoracle.fptr(manual), 512, pyoracle.fptr(s2))
    L.orc_proc_free(q)
    lim = L.orc_limiter_new(-0.5, 50.0, FS, 2.0)
    L.orc_limiter_free(lim)
    lim = L.orc_limiter_new(-1.5, 50.0, FS, 2.0)
    L.orc_limiter_process(lim, pyoracle.fptr(manual), 512)
    L.orc_limiter_free(lim)
    tpl = L.orc_tpl_new(C.c_float(FS), C.c_float(-1.5), C.c_float(80.0))
    L.orc_tpl_set_ceiling_linear(tpl, C.c_float(np.float32(10.0) ** np.float32(-1.5 / 20.0)))
    s4 = np.zeros(4, dtype=np.float32)
    L.orc_tpl_process(tpl, pyoracle.fptr(manual), 512, pyoracle.fptr(s4))
    L.orc_tpl_free(tpl)

    assert _rms_error_db(offline, manual) < -100.0
    assert np.isfinite(stats[3]) and np.isfinite(stats[2])
