"""Speech-aware auto-makeup scoring over batched control renders.

The reference scores VAD-driven auto makeup one clip at a time (python/tools/evaluate_auto_makeup_real_speech.py:
`_run_clip`, :155-273): two `simulate_auto_makeup_control` renders per clip (with the VAD evidence = candidate,
without = baseline) and a handful of statistics of the makeup traces and block boundaries.  Here all renders of
all clips go through ONE `afsim_auto_makeup_sweep` pass; the statistics are the reference's, restated
(`control_probabilities` :100-118, `block_rms_db` :121-137, `pumping_score` :140-151, clip metrics :215-273).
The VAD model itself (Silero posteriors, `analyze_vad_probabilities`) is outside the chain-simulator path: callers
pass the probabilities.
"""
from __future__ import annotations

from typing import Any, Sequence

import numpy as np

from . import abi

SAMPLE_RATE = 48_000
CONTROL_BLOCK_SIZE = abi.MAKEUP_CONTROL_BLOCK
CONTROL_CADENCE_HZ = SAMPLE_RATE / CONTROL_BLOCK_SIZE


def control_probabilities(frame_probabilities, sample_count: int, block_count: int) -> np.ndarray:
    """Frame posteriors -> one probability per 10 ms control block (linear interpolation at block centres)."""
    frame_probabilities = np.asarray(frame_probabilities, dtype=np.float64)
    if frame_probabilities.size == 0:
        return np.zeros(block_count, dtype=np.float64)
    duration = sample_count / SAMPLE_RATE
    source_times = (np.arange(frame_probabilities.size) + 0.5) * (duration / frame_probabilities.size)
    target_times = (np.arange(block_count) + 0.5) / CONTROL_CADENCE_HZ
    return np.interp(target_times, source_times, frame_probabilities, left=float(frame_probabilities[0]),
                     right=float(frame_probabilities[-1]))


def block_rms_db(audio) -> np.ndarray:
    audio = np.asarray(audio)
    out = []
    for start in range(0, audio.size, CONTROL_BLOCK_SIZE):
        block = audio[start:start + CONTROL_BLOCK_SIZE]
        out.append(20.0 * np.log10(max(float(np.sqrt(np.mean(np.square(block, dtype=np.float64)))), 1e-9)))
    return np.asarray(out, dtype=np.float64)


def pumping_score(trace_db) -> float:
    """Share of the (Hann-windowed, mean-removed) trace's spectral energy between 2 and 8 Hz."""
    trace_db = np.asarray(trace_db, dtype=np.float64)
    if trace_db.size < 10:
        return 0.0
    centered = trace_db - np.mean(trace_db)
    spectrum = np.fft.rfft(centered * np.hanning(centered.size))
    frequencies = np.fft.rfftfreq(centered.size, 1.0 / CONTROL_CADENCE_HZ)
    band = (frequencies >= 2.0) & (frequencies <= 8.0)
    total = float(np.sum(np.square(np.abs(spectrum))))
    if total <= 1e-12 or not np.any(band):
        return 0.0
    return float(np.sqrt(np.sum(np.square(np.abs(spectrum[band]))) / total))


def noise_floor_from_blocks(noisy_rms_db: np.ndarray, inactive: np.ndarray) -> float:
    """Median level of the VAD-inactive blocks, else the 20th percentile (:180-185)."""
    return float(np.median(noisy_rms_db[inactive])) if np.any(inactive) else float(np.percentile(noisy_rms_db, 20.0))


def clip_metrics(noisy, clean_control, candidate_gain, baseline_gain, candidate_output, baseline_output) -> dict[str, Any]:
    """The statistics `_run_clip` reports for one clip (:215-273), from the two renders' makeup traces and audio."""
    noisy = np.asarray(noisy, dtype=np.float64)
    clean_control = np.asarray(clean_control, dtype=np.float64)
    active, inactive = clean_control >= 0.48, clean_control <= 0.20
    candidate_gain = np.asarray(candidate_gain, dtype=np.float64)
    baseline_gain = np.asarray(baseline_gain, dtype=np.float64)
    candidate_output = np.asarray(candidate_output, dtype=np.float64)
    baseline_output = np.asarray(baseline_output, dtype=np.float64)
    count = min(candidate_gain.size, baseline_gain.size, active.size)
    active, inactive = active[:count], inactive[:count]
    candidate_gain, baseline_gain = candidate_gain[:count], baseline_gain[:count]

    def masked_median(values, mask):
        return float(np.median(values[mask])) if np.any(mask) else 0.0

    boundaries = np.arange(CONTROL_BLOCK_SIZE, noisy.size, CONTROL_BLOCK_SIZE)
    input_jumps = np.abs(noisy[boundaries] - noisy[boundaries - 1])
    candidate_excess = np.maximum(np.abs(candidate_output[boundaries] - candidate_output[boundaries - 1]) - input_jumps, 0.0)
    baseline_excess = np.maximum(np.abs(baseline_output[boundaries] - baseline_output[boundaries - 1]) - input_jumps, 0.0)
    return {
        "duration_seconds": noisy.size / SAMPLE_RATE,
        "active_block_ratio": float(np.mean(active)) if count else 0.0,
        "inactive_block_ratio": float(np.mean(inactive)) if count else 0.0,
        "candidate_active_makeup_db": masked_median(candidate_gain, active),
        "baseline_active_makeup_db": masked_median(baseline_gain, active),
        "candidate_inactive_makeup_db": masked_median(candidate_gain, inactive),
        "baseline_inactive_makeup_db": masked_median(baseline_gain, inactive),
        "candidate_pumping_score": pumping_score(candidate_gain),
        "baseline_pumping_score": pumping_score(baseline_gain),
        "candidate_max_transition_db": float(np.max(np.abs(np.diff(candidate_gain)), initial=0.0)),
        "baseline_max_transition_db": float(np.max(np.abs(np.diff(baseline_gain)), initial=0.0)),
        "candidate_p99_boundary_excess_linear": float(np.percentile(candidate_excess, 99.0)) if candidate_excess.size else 0.0,
        "baseline_p99_boundary_excess_linear": float(np.percentile(baseline_excess, 99.0)) if baseline_excess.size else 0.0,
        "candidate_max_boundary_excess_linear": float(np.max(candidate_excess, initial=0.0)),
        "candidate_final_makeup_db": float(candidate_gain[-1]) if count else 0.0,
        "baseline_final_makeup_db": float(baseline_gain[-1]) if count else 0.0,
    }


def score_clips(sim, clips: Sequence[tuple], *, settings: dict | None = None) -> list[dict[str, Any]]:
    """clips: (noisy f32 audio, clean-speech control probabilities, noisy control probabilities) per clip, 48 kHz.

    Candidate (VAD evidence) and baseline (no evidence) renders of every clip run in one GPU pass
    (`Simulator.auto_makeup_sweep`); settings default to the tool's (:187-191: vad_reliability 1, adaptive release)."""
    overrides = dict(vad_reliability=1.0, adaptive_release=True)
    overrides.update(settings or {})
    st = abi.make_makeup_settings(**overrides)
    captures, vads, floors, rels = [], [], [], []
    for noisy, clean_control, noisy_control in clips:
        noisy = np.ascontiguousarray(noisy, dtype=np.float32)
        inactive = np.asarray(clean_control, dtype=np.float64) <= 0.20
        floor = noise_floor_from_blocks(block_rms_db(noisy), inactive)
        for vad in (np.asarray(noisy_control, dtype=np.float64), None):
            captures.append(noisy)
            vads.append(vad)
            floors.append(floor)
            rels.append(1.0)
    traces, outs = sim.auto_makeup_sweep(captures, SAMPLE_RATE, vads, floors, rels, [st] * len(captures), return_audio=True)
    results = []
    for k, (noisy, clean_control, _) in enumerate(clips):
        row = clip_metrics(noisy, clean_control, traces[2 * k][0], traces[2 * k + 1][0], outs[2 * k], outs[2 * k + 1])
        row["noise_floor_db"] = floors[2 * k]
        results.append(row)
    return results
