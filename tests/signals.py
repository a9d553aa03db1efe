"""Deterministic test signals shared by the oracle and GPU parity tests."""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1


def lcg_noise(n: int, state: int = 0x6A09E667F3BCC909, scale: float = 0.012):
    """The reference golden test's LCG noise (processor/tests.rs:1812,1835-1842) -> (f64 noise, state)."""
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        state = (state * 6364136223846793005 + 1442695040888963407) & MASK64
        out[i] = (((state >> 40) & 0xFFFFFFFF) / float((1 << 24) - 1) * 2.0 - 1.0) * scale
    return out, state


def golden_chain_input(blocks: int = 300, block: int = 480, fs: float = 48000.0) -> np.ndarray:
    """Input of test_full_downstream_chain_matches_golden_tolerance (processor/tests.rs:1821-1851)."""
    n = blocks * block
    idx = np.arange(n)
    t = idx.astype(np.float64) / fs
    phrase = 0.25 + 0.75 * np.abs(np.sin(2.0 * np.pi * 1.7 * t))
    gate = (((idx // block) // 12) % 5 == 2).astype(np.float64)
    noise, _ = lcg_noise(n)
    x = (phrase * (0.30 * np.sin(2.0 * np.pi * 180.0 * t) + 0.14 * np.sin(2.0 * np.pi * 360.0 * t)
                   + 0.08 * np.sin(2.0 * np.pi * 2700.0 * t))
         + gate * 0.35 * np.sin(2.0 * np.pi * 7200.0 * t) + noise)
    return x.astype(np.float32)


def speech_like(n: int, seed: int = 0, fs: float = 48000.0, level: float = 0.5) -> np.ndarray:
    """Speech-like synthetic passage (harmonics x syllabic envelope + sibilant bursts + noise), SURVEY 8(d)."""
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
    x *= level / max(1e-9, np.max(np.abs(x)))
    return x.astype(np.float32)
