"""Batched form of the reference's Auto-EQ headroom validation.

``apply_headroom_validation`` (python/mic_eq/analysis/auto_eq_parts/headroom.py:292-354) renders the
seven gain scales one native call at a time and stops at the first safe one.  Here all scales (of
one or many EQ settings) are rendered in ONE sweep and the reference's selection rule is applied to
the finished metrics, so the outcome is identical to the sequential walk: scales are examined in
order, the first safe one wins, the last one examined is kept when none is safe, and the result
abstains (status "risk", confidences capped) when the kept render is unsafe.

Decision logic only -- the renders come from ``mic_eq_core.simulate_auto_eq_chain_batch``.
"""
from __future__ import annotations

from collections.abc import Mapping, Sequence
from copy import deepcopy
from typing import Any, Callable

import numpy as np

NUM_EQ_BANDS = 10
HEADROOM_TARGET_DB = 1.0                    # headroom.py:14
LIMITER_GAIN_REDUCTION_WARN_DB = 1.0        # headroom.py:15
TRUE_PEAK_GAIN_REDUCTION_WARN_DB = 0.5      # headroom.py:16
HEADROOM_SCALES = (1.0, 0.85, 0.70, 0.55, 0.40, 0.25, 0.0)  # headroom.py:17


def _as_float(value: Any, default: float) -> float:  # headroom.py:28-33
    try:
        parsed = float(value)
    except (TypeError, ValueError):
        return default
    return parsed if np.isfinite(parsed) else default


def _as_bool(value: Any, default: bool) -> bool:  # headroom.py:36-39
    return value if isinstance(value, bool) else default


def flatten_chain_settings(chain_settings: Mapping[str, Any] | None) -> dict[str, Any]:
    """headroom.py:42-84: nested stage settings -> the flat dict simulate_auto_eq_chain takes."""
    chain_settings = chain_settings or {}
    deesser = chain_settings.get("deesser") or {}
    compressor = chain_settings.get("compressor") or {}
    limiter = chain_settings.get("limiter") or {}
    return {
        "return_output_audio": _as_bool(chain_settings.get("return_output_audio"), False),
        "deesser_enabled": _as_bool(deesser.get("enabled"), False),
        "deesser_auto_enabled": _as_bool(deesser.get("auto_enabled"), True),
        "deesser_auto_amount": _as_float(deesser.get("auto_amount"), 0.5),
        "deesser_low_cut_hz": _as_float(deesser.get("low_cut_hz"), 4000.0),
        "deesser_high_cut_hz": _as_float(deesser.get("high_cut_hz"), 11000.0),
        "deesser_threshold_db": _as_float(deesser.get("threshold_db"), -28.0),
        "deesser_ratio": _as_float(deesser.get("ratio"), 4.0),
        "deesser_attack_ms": _as_float(deesser.get("attack_ms"), 2.0),
        "deesser_release_ms": _as_float(deesser.get("release_ms"), 80.0),
        "deesser_max_reduction_db": _as_float(deesser.get("max_reduction_db"), 6.0),
        "compressor_enabled": _as_bool(compressor.get("enabled"), True),
        "compressor_threshold_db": _as_float(compressor.get("threshold_db"), -20.0),
        "compressor_ratio": _as_float(compressor.get("ratio"), 4.0),
        "compressor_attack_ms": _as_float(compressor.get("attack_ms"), 10.0),
        "compressor_release_ms": _as_float(compressor.get("release_ms"), 200.0),
        "compressor_makeup_gain_db": _as_float(compressor.get("makeup_gain_db"), 0.0),
        "compressor_adaptive_release": _as_bool(compressor.get("adaptive_release"), False),
        "compressor_base_release_ms": _as_float(compressor.get("base_release_ms"), 50.0),
        "compressor_auto_makeup_enabled": _as_bool(compressor.get("auto_makeup_enabled"), False),
        "compressor_target_lufs": _as_float(compressor.get("target_lufs"), -18.0),
        "compressor_sidechain_highpass_enabled": _as_bool(compressor.get("sidechain_highpass_enabled"), True),
        "limiter_enabled": _as_bool(limiter.get("enabled"), True),
        "limiter_ceiling_db": _as_float(limiter.get("ceiling_db"), -0.5),
        "limiter_release_ms": _as_float(limiter.get("release_ms"), 50.0),
        "limiter_careful_output_enabled": _as_bool(limiter.get("careful_output_enabled"), True),
    }


def bands_from_settings(eq_settings: Mapping[str, Any]) -> list[tuple[float, float, float]]:
    """headroom.py:87-96."""
    freqs = list(eq_settings.get("band_freqs") or [])
    gains = list(eq_settings.get("band_gains") or [])
    qs = list(eq_settings.get("band_qs") or [])
    if not (len(freqs) == len(gains) == len(qs) == NUM_EQ_BANDS):
        raise ValueError("Auto-EQ settings must contain 10 frequencies, gains, and Q values")
    return [(_as_float(f, 1000.0), _as_float(g, 0.0), _as_float(q, 1.41)) for f, g, q in zip(freqs, gains, qs)]


def is_headroom_safe(simulation: Mapping[str, Any]) -> bool:
    """headroom.py:278-289."""
    pre = _as_float(simulation.get("pre_limiter_true_peak_headroom_db"), simulation.get("true_peak_headroom_db", 120.0))
    limiter_gr = _as_float(simulation.get("limiter_gain_reduction_db"), 0.0)
    true_peak_gr = _as_float(simulation.get("true_peak_limiter_gain_reduction_db"), 0.0)
    return (pre >= HEADROOM_TARGET_DB and limiter_gr <= LIMITER_GAIN_REDUCTION_WARN_DB
            and true_peak_gr <= TRUE_PEAK_GAIN_REDUCTION_WARN_DB)


DECISION_MARGIN_DB = 0.01  # the parity tolerance of the rendered metrics (north_star): closer than this = auditable


def headroom_margins_db(simulation: Mapping[str, Any]) -> dict[str, float]:
    """Signed distance of each of the three metrics from its threshold (>= 0: on the safe side).  A render is "near a
    threshold" when any |margin| < DECISION_MARGIN_DB: there a <= 2-ulp device-libm difference could in principle
    flip is_headroom_safe against a CPU render, so callers can audit exactly those (no CPU re-render happens here)."""
    pre = _as_float(simulation.get("pre_limiter_true_peak_headroom_db"), simulation.get("true_peak_headroom_db", 120.0))
    limiter_gr = _as_float(simulation.get("limiter_gain_reduction_db"), 0.0)
    true_peak_gr = _as_float(simulation.get("true_peak_limiter_gain_reduction_db"), 0.0)
    return {"pre_limiter_true_peak_headroom_db": pre - HEADROOM_TARGET_DB,
            "limiter_gain_reduction_db": LIMITER_GAIN_REDUCTION_WARN_DB - limiter_gr,
            "true_peak_limiter_gain_reduction_db": TRUE_PEAK_GAIN_REDUCTION_WARN_DB - true_peak_gr}


def decision_margin_report(simulations: Sequence[Mapping[str, Any]], examined: int) -> dict[str, Any]:
    """Margins of the renders the sequential walk really examined (scales 0 .. `examined`): how many sit within
    DECISION_MARGIN_DB of a threshold, and the smallest |margin| seen."""
    near, smallest, which = 0, float("inf"), None
    for index in range(examined + 1):
        margins = headroom_margins_db(simulations[index])
        closest = min(margins, key=lambda k: abs(margins[k]))
        if abs(margins[closest]) < DECISION_MARGIN_DB:
            near += 1
        if abs(margins[closest]) < smallest:
            smallest, which = abs(margins[closest]), (index, closest)
    return {"renders_examined": examined + 1, "near_threshold": near, "margin_db": DECISION_MARGIN_DB,
            "smallest_abs_margin_db": smallest, "smallest_at": {"scale_index": which[0], "metric": which[1]} if which else None}


def select_scale(simulations: Sequence[Mapping[str, Any]]) -> int:
    """Index into HEADROOM_SCALES the sequential walk of headroom.py:306-320 ends on."""
    for index, simulation in enumerate(simulations):
        if is_headroom_safe(simulation):
            return index
    return len(simulations) - 1


def scaled_candidates(eq_settings: Mapping[str, Any]) -> list[dict[str, Any]]:
    """The seven EQ settings the reference would try, in order (headroom.py:310-313)."""
    original = np.asarray(eq_settings.get("band_gains", []), dtype=float)
    out = []
    for scale in HEADROOM_SCALES:
        candidate = deepcopy(dict(eq_settings))
        candidate["band_gains"] = original.tolist() if scale == 1.0 else (original * scale).tolist()
        out.append(candidate)
    return out


def finish_validation(eq_settings: Mapping[str, Any], simulations: Sequence[Mapping[str, Any]]) -> dict[str, Any]:
    """headroom.py:322-354 applied to the seven finished renders of one EQ setting."""
    result = deepcopy(dict(eq_settings))
    original = np.asarray(result.get("band_gains", []), dtype=float)
    index = select_scale(simulations)
    selected, scale, before = simulations[index], HEADROOM_SCALES[index], simulations[0]
    result["band_gains"] = (original if index == 0 else original * scale).tolist()
    result["validation_gain_scale"] = float(_as_float(result.get("validation_gain_scale"), 1.0) * scale)
    meets = is_headroom_safe(selected)
    authoritative = selected.get("simulation_backend") == "rust"
    safe = bool(meets and authoritative)
    if not safe:
        result["validation_confidence"] = float(min(_as_float(result.get("validation_confidence"), 1.0), 0.42))
        result["analysis_confidence"] = float(min(_as_float(result.get("analysis_confidence"), 1.0), 0.58))
    elif scale < 1.0:
        result["validation_confidence"] = float(min(_as_float(result.get("validation_confidence"), 1.0), 0.72))
    result["headroom_validation"] = {
        "safe": safe, "authoritative": authoritative, "advisory": not authoritative,
        "meets_advisory_thresholds": meets, "gain_scale": scale, "before": dict(before), "after": dict(selected),
        "status": "safe" if safe else "risk" if authoritative else "advisory",
        # not part of the reference's dict: how close the examined renders came to a decision threshold
        "decision_margins": decision_margin_report(simulations, index),
    }
    result["headroom_safe"] = safe
    result["headroom_advisory"] = not authoritative
    result["headroom_gain_scale"] = scale
    return result


def apply_headroom_validation_batch(audio_data, sample_rate: float, eq_settings_list: Sequence[Mapping[str, Any]],
                                    chain_settings: Mapping[str, Any] | None = None, *,
                                    simulate_batch: Callable | None = None) -> list[dict[str, Any]]:
    """Headroom validation of many EQ settings against one capture in a single native sweep.

    Same result per setting as the reference's sequential ``apply_headroom_validation`` (settings whose
    ``band_gains`` is not 10 long are returned unchanged, headroom.py:303-304)."""
    if simulate_batch is None:
        from . import mic_eq_core
        simulate_batch = mic_eq_core.simulate_auto_eq_chain_batch
    audio = np.ascontiguousarray(np.asarray(audio_data, dtype=np.float32))
    flat = flatten_chain_settings(chain_settings)
    flat.pop("return_output_audio", None)
    jobs, owners = [], []
    for i, eq in enumerate(eq_settings_list):
        if np.asarray(eq.get("band_gains", []), dtype=float).size != NUM_EQ_BANDS:
            continue
        for candidate in scaled_candidates(eq):
            jobs.append((bands_from_settings(candidate), flat))
            owners.append(i)
    sims = simulate_batch([audio], float(sample_rate), jobs) if jobs else []
    for sim in sims:  # headroom.py:262-265
        sim["simulation_backend"] = "rust"
        sim["safety_authority"] = "authoritative"
    out: list[dict[str, Any]] = []
    cursor = 0
    for i, eq in enumerate(eq_settings_list):
        if np.asarray(eq.get("band_gains", []), dtype=float).size != NUM_EQ_BANDS:
            out.append(deepcopy(dict(eq)))
            continue
        n = len(HEADROOM_SCALES)
        out.append(finish_validation(eq, sims[cursor:cursor + n]))
        cursor += n
    return out
