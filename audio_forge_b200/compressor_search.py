"""Batched form of Auto Voice Setup's compressor calibration search.

``_calibrate_compressor_threshold`` (python/mic_eq/analysis/voice_setup.py:742-1079) evaluates up to 67
compressor candidates one native render at a time: the incumbent, 33 thresholds, 16 Halton points, then
+-local steps around two seeds; scores each render with a Huber multi-objective; and picks with a
tie-break rule that prefers a threshold-only change.  The candidate lists of both phases are known before
any render of that phase runs (dedupe and budget depend on the candidate values only), so each phase
becomes ONE sweep here -- three native calls instead of up to 68 -- and the reference's scoring,
ordering and selection are applied unchanged to the finished metrics.

Decision logic only; the renders come from ``mic_eq_core.simulate_auto_eq_chain_batch``.
"""
from __future__ import annotations

import time
from collections.abc import Callable, Mapping
from typing import Any

import numpy as np

from . import headroom

SEARCH_BUDGET = 68  # voice_setup.py:699
SEARCH_BOUNDS = {   # voice_setup.py:700-705
    "threshold_db": (-55.0, -6.0),
    "ratio": (1.5, 6.0),
    "attack_ms": (3.0, 25.0),
    "release_ms": (60.0, 320.0),
}
OBJECTIVE_NORMALIZERS = {  # voice_setup.py:706-714
    "loudness_error_db": 2.0,
    "median_gr_error_db": 1.0,
    "p95_gr_error_db": 1.0,
    "headroom_shortfall_db": 1.0,
    "pumping_score_db": 1.0,
    "silence_gain_excess_db": 1.0,
    "activity_ratio_deficit": 0.20,
}
OBJECTIVE_WEIGHTS = {  # voice_setup.py:715-724
    "loudness": 1.00,
    "median_gr": 0.35,
    "p95_gr": 0.90,
    "headroom": 0.45,
    "pumping": 0.30,
    "silence_gain": 1.50,
    "activity": 0.25,
    "prior": 0.08,
}
LOCAL_STEPS = {"threshold_db": 3.0, "ratio": 0.5, "attack_ms": 3.0, "release_ms": 25.0}  # voice_setup.py:937-942
SEARCH_LIMITER = {"enabled": True, "ceiling_db": -1.5, "release_ms": 80.0, "careful_output_enabled": True}


def _clamp(value: float, lower: float, upper: float) -> float:
    return max(lower, min(upper, value))


def huber(value: float) -> float:  # voice_setup.py:727-729
    magnitude = abs(float(value))
    return 0.5 * magnitude * magnitude if magnitude <= 1.0 else magnitude - 0.5


def halton(index: int, base: int) -> float:  # voice_setup.py:732-739
    result = 0.0
    scale = 1.0
    while index > 0:
        scale /= base
        result += scale * (index % base)
        index //= base
    return result


def key_for(candidate: Mapping[str, float]) -> tuple[float, ...]:  # voice_setup.py:779-780
    return tuple(round(float(candidate[key]), 6) for key in SEARCH_BOUNDS)


def score_simulation(simulation: Mapping[str, Any], candidate: Mapping[str, Any], incumbent: Mapping[str, float],
                     calibrated: Mapping[str, Any], target_p95_db: float, target_median_db: float,
                     peak_cap_db: float) -> float:
    """voice_setup.py:817-909: Huber multi-objective of one finished render (inf when hard-rejected)."""
    if simulation.get("simulation_backend") != "rust":
        return float("inf")
    peak = float(simulation.get("compressor_gain_reduction_db", 0.0))
    median = float(simulation.get("compressor_gain_reduction_median_db", peak))
    p95 = float(simulation.get("compressor_gain_reduction_p95_db", peak))
    active_ratio = float(simulation.get("compressor_gain_reduction_active_ratio", 0.0))
    active_gain = float(simulation.get("active_output_gain_db", 0.0))
    target_lufs = float(calibrated.get("target_lufs", -18.0))
    output_lufs = (target_lufs if calibrated.get("auto_makeup_enabled", False)
                   else float(calibrated.get("measured_short_term_lufs", -18.0)) + active_gain)
    output_true_peak = float(simulation.get("output_true_peak_db", 120.0))
    ceiling = float(simulation.get("limiter_effective_ceiling_db", -1.5))
    pre_limiter_headroom = float(simulation.get("pre_limiter_true_peak_headroom_db", -120.0))
    pumping = float(simulation.get("compressor_pumping_score_db", 120.0))
    silence_gain = float(simulation.get("silence_output_gain_db", 120.0))
    non_finite = bool(simulation.get("non_finite_output", True))
    finite_values = np.asarray([peak, median, p95, active_ratio, output_lufs, output_true_peak, pre_limiter_headroom,
                                pumping, silence_gain], dtype=float)
    hard_rejected = bool(non_finite or not np.isfinite(finite_values).all() or output_true_peak > ceiling + 0.10
                         or peak > peak_cap_db + 1.0e-6)
    prior_terms = []
    for key, (lower, upper) in SEARCH_BOUNDS.items():
        span = upper - lower
        prior_terms.append(((float(candidate[key]) - incumbent[key]) / span) ** 2)
    terms = {
        "loudness": huber((output_lufs - target_lufs) / OBJECTIVE_NORMALIZERS["loudness_error_db"]),
        "median_gr": huber((median - target_median_db) / OBJECTIVE_NORMALIZERS["median_gr_error_db"]),
        "p95_gr": huber((p95 - target_p95_db) / OBJECTIVE_NORMALIZERS["p95_gr_error_db"]),
        "headroom": huber(max(0.0, 1.0 - pre_limiter_headroom) / OBJECTIVE_NORMALIZERS["headroom_shortfall_db"]),
        "pumping": huber(pumping / OBJECTIVE_NORMALIZERS["pumping_score_db"]),
        "silence_gain": huber(max(0.0, silence_gain - 0.25) / OBJECTIVE_NORMALIZERS["silence_gain_excess_db"]),
        "activity": huber(max(0.0, 0.20 - active_ratio) / OBJECTIVE_NORMALIZERS["activity_ratio_deficit"]),
        "prior": float(np.mean(prior_terms)),
    }
    score = sum(OBJECTIVE_WEIGHTS[name] * value for name, value in terms.items())
    return float("inf") if hard_rejected else float(score)


def select_winner(evaluated: Mapping[tuple, tuple[float, Mapping, Mapping[str, float]]], incumbent: Mapping[str, float]):
    """voice_setup.py:968-995 -> (best entry, expanded entry, threshold-only entry | None, expanded_selected)."""
    feasible = sorted((item for item in evaluated.values() if np.isfinite(item[0])), key=lambda item: (item[0], key_for(item[2])))
    threshold_only = min(
        (item for item in feasible
         if all(abs(item[2][key] - incumbent[key]) <= 1.0e-6 for key in ("ratio", "attack_ms", "release_ms"))),
        key=lambda item: (item[0], key_for(item[2])), default=None)
    expanded = feasible[0]
    if threshold_only is None:
        return expanded, expanded, None, True
    required = max(0.001, 0.01 * threshold_only[0])
    expanded_selected = bool(threshold_only[0] - expanded[0] > required)
    return (expanded if expanded_selected else threshold_only), expanded, threshold_only, expanded_selected


TIE_MARGIN = 1.0e-5          # score distance below which the argmin / the expanded-vs-threshold-only rule is "near"
HARD_REJECT_MARGIN_DB = 0.01  # the parity tolerance of the rendered metrics


def search_decision_margins(evaluated: Mapping[tuple, tuple[float, Mapping, Mapping[str, float]]], incumbent: Mapping[str, float],
                            peak_cap_db: float) -> dict[str, Any]:
    """How close the search's decisions came to flipping (no CPU re-render; counts only, so that "bit-exact
    selection" is auditable): hard-reject tests within HARD_REJECT_MARGIN_DB of their limit (voice_setup.py:862-867),
    the gap between the two best scores, and the distance of the expanded-vs-threshold-only rule (voice_setup.py:
    984-995) from its required improvement."""
    near_reject = 0
    for _, sim, _ in evaluated.values():
        ceiling = float(sim.get("limiter_effective_ceiling_db", -1.5))
        d_peak = abs(float(sim.get("output_true_peak_db", 120.0)) - (ceiling + 0.10))
        d_cap = abs(float(sim.get("compressor_gain_reduction_db", 0.0)) - (peak_cap_db + 1.0e-6))
        if min(d_peak, d_cap) < HARD_REJECT_MARGIN_DB:
            near_reject += 1
    feasible = sorted(item[0] for item in evaluated.values() if np.isfinite(item[0]))
    best_gap = feasible[1] - feasible[0] if len(feasible) > 1 else float("inf")
    _, expanded, threshold_only, _ = select_winner(evaluated, incumbent) if feasible else (None, None, None, False)
    rule_distance = float("inf")
    if threshold_only is not None and expanded is not None:
        required = max(0.001, 0.01 * threshold_only[0])
        rule_distance = abs((threshold_only[0] - expanded[0]) - required)
    return {"candidates": len(evaluated), "near_hard_reject": near_reject, "hard_reject_margin_db": HARD_REJECT_MARGIN_DB,
            "best_score_gap": best_gap, "tie_rule_distance": rule_distance, "tie_margin": TIE_MARGIN,
            "near_tie": bool(best_gap < TIE_MARGIN or rule_distance < TIE_MARGIN)}


def calibrate_compressor_batch(*, speech_audio, sample_rate: float, eq_settings: Mapping[str, Any],
                               deesser_settings: Mapping[str, Any], compressor_settings: Mapping[str, Any],
                               target_p95_db: float, target_median_db: float, peak_cap_db: float,
                               simulate_batch: Callable | None = None) -> tuple[dict[str, Any], dict[str, Any]]:
    """Same result as the reference's sequential search, with one native sweep per search phase."""
    if simulate_batch is None:
        from . import mic_eq_core
        simulate_batch = mic_eq_core.simulate_auto_eq_chain_batch
    started = time.perf_counter()
    audio = np.ascontiguousarray(np.asarray(speech_audio, dtype=np.float32))
    bands = headroom.bands_from_settings(eq_settings)
    calibrated = dict(compressor_settings)
    incumbent = {key: _clamp(float(calibrated[key]), *SEARCH_BOUNDS[key]) for key in SEARCH_BOUNDS}
    evaluated: dict[tuple, tuple[float, dict, dict[str, float]]] = {}
    native_calls = 0

    def run_phase(candidate_values_list):
        """The reference's evaluate() (voice_setup.py:782-914) over a list: same dedupe, same budget, one sweep."""
        nonlocal native_calls
        pending, keys = [], []
        for values in candidate_values_list:
            if len(evaluated) + len(pending) >= SEARCH_BUDGET - 1:
                break
            k = key_for(values)
            if k in evaluated or k in keys:
                continue
            candidate = dict(calibrated)
            candidate.update({key: _clamp(float(values[key]), *SEARCH_BOUNDS[key]) for key in SEARCH_BOUNDS})
            simulation_compressor = dict(candidate)
            if simulation_compressor.get("auto_makeup_enabled", False):
                simulation_compressor["auto_makeup_enabled"] = False
                simulation_compressor["makeup_gain_db"] = 0.0
            flat = headroom.flatten_chain_settings({"deesser": deesser_settings, "compressor": simulation_compressor,
                                                    "limiter": SEARCH_LIMITER})
            flat.pop("return_output_audio", None)
            pending.append((candidate, flat))
            keys.append(k)
        if not pending:
            return
        sims = simulate_batch([audio], float(sample_rate), [(bands, flat) for _, flat in pending])
        native_calls += 1
        for k, (candidate, _), sim in zip(keys, pending, sims):
            sim = dict(sim)
            sim["simulation_backend"] = "rust"            # headroom.py:262-265
            sim["safety_authority"] = "authoritative"
            score = score_simulation(sim, candidate, incumbent, calibrated, target_p95_db, target_median_db, peak_cap_db)
            evaluated[k] = (score, sim, {key: float(candidate[key]) for key in SEARCH_BOUNDS})

    # phase 1 (voice_setup.py:916-926): incumbent, 33 thresholds, 16 Halton points
    phase1 = [dict(incumbent)]
    for threshold in np.linspace(-55.0, -6.0, 33):
        c = dict(incumbent)
        c["threshold_db"] = float(threshold)
        phase1.append(c)
    for index in range(1, 17):
        c = {}
        for key, base in zip(SEARCH_BOUNDS, (2, 3, 5, 7)):
            lower, upper = SEARCH_BOUNDS[key]
            c[key] = lower + halton(index, base) * (upper - lower)
        phase1.append(c)
    run_phase(phase1)

    diagnostics: dict[str, Any] = {
        "backend": "rust", "objective": "bounded_multi_objective_compressor_search_v1",
        "target_p95_gain_reduction_db": target_p95_db, "target_median_gain_reduction_db": target_median_db,
        "peak_gain_reduction_cap_db": peak_cap_db, "candidate_budget": SEARCH_BUDGET,
    }
    feasible = sorted((item for item in evaluated.values() if np.isfinite(item[0])), key=lambda item: (item[0], key_for(item[2])))
    if not feasible:  # voice_setup.py:932-935
        diagnostics["iterations"] = len(evaluated)
        diagnostics["native_calls"] = native_calls
        diagnostics["search_runtime_ms"] = (time.perf_counter() - started) * 1000.0
        return calibrated, diagnostics

    # phase 2 (voice_setup.py:943-966): +-local steps around the best and a multivariable seed
    seeds = [feasible[0]]
    multivariable = next((item for item in feasible
                          if any(abs(item[2][key] - incumbent[key]) > 1.0e-6 for key in ("ratio", "attack_ms", "release_ms"))),
                         None)
    if multivariable is not None and key_for(multivariable[2]) != key_for(seeds[0][2]):
        seeds.append(multivariable)
    else:
        seeds.extend(feasible[1:2])
    phase2 = []
    for _, _, seed in seeds:
        for key, step in LOCAL_STEPS.items():
            for direction in (-1.0, 1.0):
                c = dict(seed)
                c[key] += direction * step
                phase2.append(c)
    run_phase(phase2)

    (best_score, best_simulation, best_values), expanded, threshold_only, expanded_selected = select_winner(evaluated, incumbent)
    calibrated.update(best_values)
    # winner verification render (voice_setup.py:997-1023)
    winner = dict(calibrated)
    if calibrated.get("auto_makeup_enabled", False):
        winner.update({"auto_makeup_enabled": False, "makeup_gain_db": 0.0})
    flat = headroom.flatten_chain_settings({"deesser": deesser_settings, "compressor": winner, "limiter": SEARCH_LIMITER})
    flat.pop("return_output_audio", None)
    verification = simulate_batch([audio], float(sample_rate), [(bands, flat)])[0]
    native_calls += 1
    best_simulation = dict(verification)
    threshold_only_scores = [score for score, _, values in evaluated.values()
                             if all(abs(values[key] - incumbent[key]) <= 1.0e-6 for key in ("ratio", "attack_ms", "release_ms"))]
    incumbent_entry = evaluated.get(key_for(incumbent))
    peak = float(best_simulation["compressor_gain_reduction_db"])
    diagnostics.update({
        "measured_median_gain_reduction_db": float(best_simulation["compressor_gain_reduction_median_db"]),
        "measured_p95_gain_reduction_db": float(best_simulation["compressor_gain_reduction_p95_db"]),
        "measured_peak_gain_reduction_db": peak,
        "active_reduction_ratio": float(best_simulation["compressor_gain_reduction_active_ratio"]),
        "peak_cap_passed": peak <= peak_cap_db + 1.0e-6,
        "total_objective": best_score,
        "incumbent_objective": incumbent_entry[0] if incumbent_entry is not None else float("inf"),
        "threshold_only_objective": min(threshold_only_scores, default=float("inf")),
        "expanded_candidate_objective": expanded[0],
        "expanded_search_selected": expanded_selected,
        "candidate_count": len(evaluated) + 1,
        "iterations": len(evaluated) + 1,
        "native_calls": native_calls,
        "search_runtime_ms": (time.perf_counter() - started) * 1000.0,
        "threshold_db": calibrated["threshold_db"], "ratio": calibrated["ratio"],
        "attack_ms": calibrated["attack_ms"], "release_ms": calibrated["release_ms"],
        # not part of the reference's diagnostics: how close the decisions came to flipping
        "decision_margins": search_decision_margins(evaluated, incumbent, peak_cap_db),
    })
    return calibrated, diagnostics
