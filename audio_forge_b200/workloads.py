"""Synthetic workloads of the named benchmark shapes (BASELINE.json configs).

Data generation only -- no DSP.  Deterministic given the seed so that every rank / arm builds the
same inputs.
"""
from __future__ import annotations

import numpy as np

from . import abi

FS = 48000.0
HEADROOM_SCALES = (1.0, 0.85, 0.70, 0.55, 0.40, 0.25, 0.0)  # headroom.py:17


def speech_like(n: int, seed: int = 0, fs: float = FS, level: float = 0.5) -> np.ndarray:
    """Harmonics of a jittered f0 x syllabic envelope + gated 7.2 kHz sibilant bursts + noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / fs
    f0 = 110.0 + 110.0 * rng.random()
    env = 0.25 + 0.75 * np.abs(np.sin(2.0 * np.pi * 1.7 * t + rng.random()))
    x = np.zeros(n)
    for k, a in enumerate((0.30, 0.14, 0.10, 0.08, 0.05, 0.03), start=1):
        x += a * np.sin(2.0 * np.pi * f0 * k * t + rng.random() * 6.28)
    x *= env
    gate = ((np.floor(t / 0.12).astype(np.int64) % 5) == 2).astype(np.float64)
    x += gate * 0.30 * np.sin(2.0 * np.pi * 7200.0 * t) * (0.5 + 0.5 * rng.random())
    x += rng.standard_normal(n) * 0.0126
    x *= level / max(1e-9, float(np.max(np.abs(x))))
    return x.astype(np.float32)


def _set_band(band: abi.AfBand, name: str, f: float, g: float, q: float, slope: int = 12, enabled: bool = True) -> None:
    band.frequency_hz, band.gain_db, band.q = float(f), float(g), float(q)
    band.filter_type = abi.FILTER_IDS[name]
    band.slope_db_per_octave = slope
    band.enabled = 1 if enabled else 0


def headroom_candidates(n_candidates: int = 4096, seed: int = 1234, **settings_overrides):
    """C2 (Auto-EQ headroom validation): base 10-band typed settings x the 7 headroom scales, +flat.

    Gains U[-12, 12] dB, Q U[0.3, 6]; in 25 % of the base settings band 0 / band 9 become 48 dB/oct
    high-pass / low-pass cascades.  Chain settings are the reference defaults (compressor on, limiter
    careful) unless overridden."""
    rng = np.random.default_rng(seed)
    cands = (abi.AfCandidate * n_candidates)()
    settings = abi.make_settings(use_typed_bands=True, **settings_overrides)
    names = ["low_shelf"] + ["bell"] * 8 + ["high_shelf"]
    i = 0
    while i < n_candidates:
        gains = rng.uniform(-12.0, 12.0, size=10)
        qs = rng.uniform(0.3, 6.0, size=10)
        passes = rng.random() < 0.25
        freqs = np.array(abi.DEFAULT_FREQUENCIES) * rng.uniform(0.9, 1.1, size=10)
        for scale in HEADROOM_SCALES:
            if i >= n_candidates:
                break
            c = cands[i]
            for b in range(10):
                if passes and b == 0:
                    _set_band(c.bands[b], "high_pass", 60.0 + 40.0 * rng.random(), 0.0, 0.707, 48)
                elif passes and b == 9:
                    _set_band(c.bands[b], "low_pass", 15000.0 + 3000.0 * rng.random(), 0.0, 0.707, 48)
                else:
                    _set_band(c.bands[b], names[b], freqs[b], gains[b] * scale, qs[b])
            c.settings = settings
            i += 1
    return cands


def compressor_grid_candidates(n_thr: int = 16, n_ratio: int = 16, n_attack: int = 8, n_release: int = 8):
    """C3 (compressor calibration grid) over the search bounds of voice_setup.py:700-705."""
    total = n_thr * n_ratio * n_attack * n_release
    cands = (abi.AfCandidate * total)()
    bands = abi.default_bands()
    i = 0
    for thr in np.linspace(-55.0, -6.0, n_thr):
        for ratio in np.linspace(1.5, 6.0, n_ratio):
            for attack in np.linspace(3.0, 25.0, n_attack):
                for release in np.linspace(60.0, 320.0, n_release):
                    c = cands[i]
                    for b in range(10):
                        c.bands[b] = bands[b]
                    c.settings = abi.make_settings(compressor_threshold_db=thr, compressor_ratio=ratio,
                                                   compressor_attack_ms=attack, compressor_release_ms=release,
                                                   compressor_adaptive_release=True, limiter_ceiling_db=-1.5,
                                                   limiter_release_ms=80.0)
                    i += 1
    return cands


def is_headroom_safe(m: dict) -> bool:
    """headroom.py:278-289."""
    return (m["pre_limiter_true_peak_headroom_db"] >= 1.0 and m["limiter_gain_reduction_db"] <= 1.0
            and m["true_peak_limiter_gain_reduction_db"] <= 0.5)
