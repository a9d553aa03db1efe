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


def _grid_shape(n: int) -> tuple[int, int, int, int]:
    """Threshold x ratio x attack x release counts whose product is n (n a power of two): 16384 -> 16x16x8x8."""
    dims = [1, 1, 1, 1]
    k = 0
    while dims[0] * dims[1] * dims[2] * dims[3] < n:
        dims[k % 4] *= 2
        k += 1
    return tuple(dims)  # type: ignore[return-value]


def compressor_grid_candidates(n_candidates: int = 16384, seed: int = 0):
    """C3 (compressor calibration grid) over the search bounds of voice_setup.py:700-705: threshold [-55,-6] dB,
    ratio [1.5,6], attack [3,25] ms, release [60,320] ms; adaptive release on; limiter as the search renders it
    (-1.5 dB, 80 ms, careful; voice_setup.py:802-816); auto-makeup off (voice_setup.py:798-801)."""
    n_thr, n_ratio, n_attack, n_release = _grid_shape(n_candidates)
    cands = (abi.AfCandidate * n_candidates)()
    bands = abi.default_bands()
    rng = np.random.default_rng(seed)
    jitter = rng.uniform(-0.01, 0.01)
    i = 0
    for thr in np.linspace(-55.0, -6.0, n_thr):
        for ratio in np.linspace(1.5, 6.0, n_ratio):
            for attack in np.linspace(3.0, 25.0, n_attack):
                for release in np.linspace(60.0, 320.0, n_release):
                    if i >= n_candidates:
                        break
                    c = cands[i]
                    for b in range(10):
                        c.bands[b] = bands[b]
                    c.settings = abi.make_settings(compressor_threshold_db=float(thr) + jitter, compressor_ratio=float(ratio),
                                                   compressor_attack_ms=float(attack), compressor_release_ms=float(release),
                                                   compressor_adaptive_release=True, limiter_ceiling_db=-1.5,
                                                   limiter_release_ms=80.0)
                    i += 1
    return cands


def default_preset_candidates(n_candidates: int = 1):
    """C1 (default preset chain render): flat legacy EQ (with its 72-sample fade-in), compressor -20 dB / 4:1 /
    10 ms / 200 ms (base release 50 ms, sidechain high-pass on), limiter -0.5 dB careful (-> -1.5 dB), de-esser
    off (config_parts/settings.py:551-591), behind the live loop's DC block + 80 Hz high-pass."""
    cands = (abi.AfCandidate * n_candidates)()
    bands = abi.default_bands()
    for i in range(n_candidates):
        for b in range(10):
            cands[i].bands[b] = bands[b]
        cands[i].settings = abi.make_settings(input_stage="dc_hp80")
    return cands


def true_peak_candidates(n_candidates: int = 1):
    """C4 (batch true-peak detection + lookahead limiting): flat typed EQ, compressor off, limiter -1.5 dB /
    50 ms / 2 ms lookahead + true-peak limiter + detector."""
    cands = (abi.AfCandidate * n_candidates)()
    bands = abi.default_bands()
    for i in range(n_candidates):
        for b in range(10):
            cands[i].bands[b] = bands[b]
        cands[i].settings = abi.make_settings(use_typed_bands=True, compressor_enabled=False, limiter_ceiling_db=-1.5,
                                              limiter_release_ms=50.0, limiter_lookahead_ms=2.0)
    return cands


def full_chain_candidates(n_candidates: int = 8192, seed: int = 0):
    """C5 (full chain): adaptive input cleanup (49-61 Hz hum / harmonic notches, Strong), auto de-esser, typed EQ,
    compressor, limiter, true peak."""
    rng = np.random.default_rng(seed)
    cands = headroom_candidates(n_candidates, seed=seed + 7)
    for i in range(n_candidates):
        cands[i].settings = abi.make_settings(
            use_typed_bands=True, input_stage=3, deesser_enabled=True, deesser_auto_enabled=True,
            deesser_auto_amount=float(rng.uniform(0.2, 0.9)), deesser_max_reduction_db=float(rng.uniform(4.0, 10.0)),
            compressor_threshold_db=float(rng.uniform(-35.0, -12.0)), compressor_ratio=float(rng.uniform(2.0, 6.0)),
            compressor_adaptive_release=bool(i % 2))
    return cands


def is_headroom_safe(m: dict) -> bool:
    """headroom.py:278-289."""
    return (m["pre_limiter_true_peak_headroom_db"] >= 1.0 and m["limiter_gain_reduction_db"] <= 1.0
            and m["true_peak_limiter_gain_reduction_db"] <= 0.5)


def add_hum(x: np.ndarray, hum_hz: float = 50.37, level_db: float = -26.0, fs: float = FS) -> np.ndarray:
    """Mains hum + second harmonic injected at level_db dBFS."""
    t = np.arange(x.size, dtype=np.float64) / fs
    a = 10.0 ** (level_db / 20.0)
    y = x.astype(np.float64) + a * np.sin(2.0 * np.pi * hum_hz * t) + 0.5 * a * np.sin(2.0 * np.pi * 2.0 * hum_hz * t + 0.3)
    return y.astype(np.float32)


def synthetic_noise_host(passage: int, n: int) -> np.ndarray:
    """Host copy of the device generator `k_synth` kind 1 (hot white noise, csrc/afsim_kernels.cu): a counter
    hash per (passage, sample), so any stream of a device-generated batch can be rebuilt for the oracle."""
    mask = (1 << 64) - 1
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (np.uint64(0x6A09E667F3BCC909) + np.uint64((passage * 0x9E3779B97F4A7C15) & mask)
             + idx * np.uint64(0xBF58476D1CE4E5B9))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    bits = ((z >> np.uint64(40)) & np.uint64(0xFFFFFF)).astype(np.float32)
    u = bits / np.float32(16777215.0) * np.float32(2.0) - np.float32(1.0)
    return (u * np.float32(0.88)).astype(np.float32)
