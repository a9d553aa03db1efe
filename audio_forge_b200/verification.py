"""Verification renders of Auto Voice Setup in one native call (first half).

``validate_voice_setup_verification`` (python/mic_eq/analysis/voice_setup.py:1468-1660) pushes the second speech
passage AND the noise capture through the exact candidate chain -- two sequential native calls with
``return_output_audio`` (:1497-1524) -- before its spectral checks.  ``render_verification_pair`` renders both in
ONE sweep (two streams of different length: two batches of one stream each, one launch) and returns what the two
``simulate_candidate_chain`` calls return, so the reference's checks downstream run unchanged on the same values.

Host logic only -- the renders come from ``mic_eq_core.simulate_auto_eq_chain_batch``; the spectral analysis
(`analyze_voice_spectrum`, `_shape_error_db`, K-weighted windows) stays with the reference's Python.
"""
from __future__ import annotations

from collections.abc import Mapping
from typing import Any, Callable

import numpy as np

from .abi import DEFAULT_FREQUENCIES
from .headroom import bands_from_settings, flatten_chain_settings

SPEECH_MIN_DURATION_S = 3.0  # voice_setup.py:43


def verification_chain(setup_result: Mapping[str, Any]) -> tuple[dict[str, Any], dict[str, Any]]:
    """(eq_settings, chain_settings) of the verification renders (voice_setup.py:1497-1515): the setup's EQ (flat
    when it has none), its de-esser and compressor, the fixed -1.5 dB / 80 ms careful limiter, audio returned."""
    eq_settings = dict(setup_result.get("eq_settings") or {})
    if not eq_settings:
        eq_settings = {"band_freqs": list(DEFAULT_FREQUENCIES), "band_gains": [0.0] * len(DEFAULT_FREQUENCIES),
                       "band_qs": [1.41] * len(DEFAULT_FREQUENCIES)}
    chain = {
        "deesser": dict(setup_result.get("deesser_settings") or {}),
        "compressor": dict(setup_result.get("compressor_settings") or {}),
        "limiter": {"enabled": True, "ceiling_db": -1.5, "release_ms": 80.0, "careful_output_enabled": True},
        "return_output_audio": True,
    }
    return eq_settings, chain


def passage_gate(verification_audio: np.ndarray, sample_rate: float) -> dict[str, Any] | None:
    """The two early exits of voice_setup.py:1484-1495 (too short; non-finite or clipped), or None."""
    verification = np.asarray(verification_audio, dtype=np.float32)
    if verification.size < int(sample_rate * SPEECH_MIN_DURATION_S):
        return {"decision": "retry", "reasons": ["verification passage was too short"], "perceptual_validation": False}
    if not np.isfinite(verification).all() or float(np.max(np.abs(verification))) >= 0.999:
        return {"decision": "retry", "reasons": ["verification passage was non-finite or clipped"],
                "perceptual_validation": False}
    return None


def render_verification_pair(noise_audio, verification_speech_audio, sample_rate: float, setup_result: Mapping[str, Any],
                             *, simulate_batch: Callable | None = None):
    """-> (processed, rendered, processed_noise, rendered_noise): the two result dicts of voice_setup.py:1516-1527
    with ``output_audio`` popped into float32 arrays (:1535-1536), from one native sweep."""
    if simulate_batch is None:
        from . import mic_eq_core
        simulate_batch = mic_eq_core.simulate_auto_eq_chain_batch
    noise = np.ascontiguousarray(np.asarray(noise_audio, dtype=np.float32))
    verification = np.ascontiguousarray(np.asarray(verification_speech_audio, dtype=np.float32))
    eq_settings, chain = verification_chain(setup_result)
    flat = flatten_chain_settings(chain)
    flat.pop("return_output_audio", None)
    job = (bands_from_settings(eq_settings), flat)
    sims = simulate_batch([verification, noise], float(sample_rate), [job], return_output_audio=True)
    out = []
    for sim in sims:  # headroom.py:262-265
        sim = dict(sim)
        sim["simulation_backend"] = "rust"
        sim["safety_authority"] = "authoritative"
        audio = np.asarray(sim.pop("output_audio"), dtype=np.float32)
        out.extend([sim, audio])
    return tuple(out)
