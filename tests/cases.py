"""Chain configurations shared by the CPU (hostsim vs oracle) and GPU (libafsim vs oracle) parity tests."""
from __future__ import annotations

import numpy as np

from audio_forge_b200 import abi

FS = 48000.0

LEGACY_EQ = [(80, 3, 1), (160, -2, 1.2), (320, 1, 1.41), (640, -4, 2), (1280, 2, 0.7), (2500, 5, 1), (5000, -6, 3),
             (8000, 4, 1), (12000, 2, 1), (16000, -3, 0.8)]
GOLDEN_EQ = [(80, 0, 1.41), (160, 0, 1.41), (180, -2.5, 0.8), (640, 0, 1.41), (1280, 0, 1.41), (2500, 0, 1.41),
             (2800, 3, 1.2), (8000, 0, 1.41), (7200, 1.5, 1.0), (16000, 0, 1.41)]
TYPED_PASS = [("high_pass", 90, 0, 0.7, 48, True), ("bell", 160, 2, 1, 12, True), ("notch", 320, 0, 4, 12, True),
              ("bell", 640, -3, 1, 12, False), ("low_shelf", 200, 3, 0.9, 12, True), ("bell", 2500, 4, 1, 12, True),
              ("bell", 5000, -5, 2, 12, True), ("high_shelf", 8000, 3, 0.8, 12, True), ("bell", 12000, 1, 1, 12, True),
              ("low_pass", 15000, 0, 0.7, 36, True)]
TYPED_WORST = [("high_pass" if i % 2 == 0 else "low_pass", 60.0 + 30 * i if i % 2 == 0 else 18000.0 - 400 * i, 0, 0.7, 48,
                True) for i in range(10)]

# name -> (bands, settings overrides)
CASES = {
    "default_legacy": (abi.default_bands(), dict()),
    "legacy_eq": (abi.legacy_bands(LEGACY_EQ), dict(compressor_makeup_gain_db=6.0)),
    # the reference's golden-vector settings (processor/tests.rs:1786-1810) through the chain-sim door
    "golden_like": (abi.legacy_bands(GOLDEN_EQ),
                    dict(deesser_enabled=True, deesser_auto_amount=0.85, deesser_max_reduction_db=10.0,
                         compressor_threshold_db=-22.0, compressor_ratio=3.5, compressor_attack_ms=8.0,
                         compressor_release_ms=160.0, compressor_makeup_gain_db=8.0, compressor_adaptive_release=True,
                         limiter_ceiling_db=-6.0, limiter_release_ms=55.0, limiter_careful_output_enabled=False)),
    "typed_pass": (abi.typed_bands(TYPED_PASS),
                   dict(use_typed_bands=True, deesser_enabled=True, deesser_auto_enabled=False, deesser_low_cut_hz=5000.0,
                        deesser_threshold_db=-40.0, eq_before_deesser=True, compressor_sidechain_highpass_enabled=False,
                        limiter_lookahead_ms=5.0)),
    "typed_worst_40_sections": (abi.typed_bands(TYPED_WORST), dict(use_typed_bands=True, compressor_makeup_gain_db=12.0)),
    "no_limiter": (abi.default_bands(), dict(limiter_enabled=False, compressor_enabled=False)),
    "dc_hp": (abi.default_bands(), dict(input_stage=1, compressor_makeup_gain_db=10.0)),
    "short_lookahead": (abi.legacy_bands(LEGACY_EQ), dict(limiter_lookahead_ms=0.1, compressor_makeup_gain_db=9.0)),
    "long_lookahead": (abi.legacy_bands(LEGACY_EQ), dict(limiter_lookahead_ms=10.0, compressor_makeup_gain_db=9.0,
                                                         limiter_ceiling_db=-3.0)),
}


def candidate(bands, **settings) -> abi.AfCandidate:
    c = abi.AfCandidate()
    for i in range(abi.NUM_BANDS):
        c.bands[i] = bands[i]
    c.settings = abi.make_settings(**settings)
    return c


def candidate_array(items):
    arr = (abi.AfCandidate * len(items))()
    for i, c in enumerate(items):
        arr[i] = c
    return arr


DISCRETE_KEYS = ("true_peak_limited_events", "non_finite_output", "active_analysis_block_count", "processed_samples")


def metric_mismatches(expected: abi.AfChainMetrics, got: abi.AfChainMetrics, tol_db: float = 0.0, exact_discrete: bool = True):
    """Keys whose values differ by more than tol_db (f32 metrics) / at all (counts)."""
    e, g = abi.metrics_to_dict(expected), abi.metrics_to_dict(got)
    bad = {}
    for key in abi.METRIC_F32_KEYS:
        a, b = e[key], g[key]
        if np.isnan(a) and np.isnan(b):
            continue
        if np.isinf(a) or np.isinf(b):
            if a != b:
                bad[key] = (a, b)
            continue
        if abs(a - b) > tol_db:
            bad[key] = (a, b)
    if exact_discrete:
        for key in DISCRETE_KEYS:
            if e[key] != g[key]:
                bad[key] = (e[key], g[key])
    return bad


def audio_within_tolerance(expected: np.ndarray, got: np.ndarray) -> float:
    """north_star tolerance: 1e-5 relative or -100 dBFS (1e-5) absolute.  Returns the worst excess (<= 0 passes)."""
    expected = expected.astype(np.float64)
    got = got.astype(np.float64)
    err = np.abs(expected - got)
    allowed = np.maximum(1e-5 * np.abs(expected), 1e-5)
    return float(np.max(err - allowed)) if err.size else 0.0
