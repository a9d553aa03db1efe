"""Host-side mirror of the reference's PyO3 module for the chain-simulator path.

Same function names, argument meaning and error behaviour as ``mic_eq.mic_eq_core``
(rust-core/src/lib.rs:99-288, rust-core/src/audio/processor/python_api.rs:378-714); every call
goes through the C ABI of ``libafsim.so`` onto the GPU.  Stands in for the Rust host in this
repository (no Rust toolchain in the image): ``mic_eq/__init__.py:38-46`` falls back to a
top-level ``import mic_eq_core``, so putting this package directory on ``sys.path`` makes the
unmodified ``mic_eq`` callers (``headroom._native_simulate``) use it.  INTEGRATION.md shows the
``extern "C"`` binding the Rust host would add instead.

Only the chain-simulator entry points exist here; live-engine names (``AudioProcessor`` ...) are
stubs that raise, exactly like a core built without them would.
"""
from __future__ import annotations

import threading
from collections.abc import Mapping, Sequence
from typing import Any

import numpy as np

try:  # imported as audio_forge_b200.mic_eq_core
    from . import abi, native
except ImportError:  # imported as top-level mic_eq_core
    from audio_forge_b200 import abi, native

NUM_BANDS = abi.NUM_BANDS
_lock = threading.Lock()
_sims: dict[int, native.Simulator] = {}


def simulator(device: int = 0) -> native.Simulator:
    """Process-wide handle per device (created on first use; raises without an sm_100 GPU)."""
    with _lock:
        sim = _sims.get(device)
        if sim is None:
            sim = native.Simulator(device)
            _sims[device] = sim
        return sim


# ---- settings dict -> AfChainSettings (python_api.rs:14-52, 415-487) -------------------------------------
_BOOL_KEYS = {
    "eq_before_deesser", "deesser_enabled", "deesser_auto_enabled", "compressor_enabled",
    "compressor_adaptive_release", "compressor_auto_makeup_enabled", "compressor_sidechain_highpass_enabled",
    "limiter_enabled", "limiter_careful_output_enabled",
}
_F64_KEYS = {
    "deesser_auto_amount", "deesser_low_cut_hz", "deesser_high_cut_hz", "deesser_threshold_db", "deesser_ratio",
    "deesser_attack_ms", "deesser_release_ms", "deesser_max_reduction_db", "compressor_threshold_db",
    "compressor_ratio", "compressor_attack_ms", "compressor_release_ms", "compressor_makeup_gain_db",
    "compressor_base_release_ms", "compressor_target_lufs", "limiter_ceiling_db", "limiter_release_ms",
    "limiter_lookahead_ms",
}


def _extract_bool(key: str, value: object) -> bool:
    if isinstance(value, (bool, np.bool_)):  # PyO3 `extract::<bool>` accepts bool only
        return bool(value)
    raise TypeError(f"settings[{key!r}]: argument must be bool, not {type(value).__name__}")


def _extract_f64(key: str, value: object) -> float:
    if isinstance(value, (bool, np.bool_)):
        return float(value)
    try:
        return float(value)  # type: ignore[arg-type]
    except (TypeError, ValueError) as error:
        raise TypeError(f"settings[{key!r}]: argument must be a real number") from error


def settings_from_mapping(settings: Mapping[str, object] | None) -> tuple[abi.AfChainSettings, Any, bool]:
    """-> (AfChainSettings, eq_bands_v2 | None, return_output_audio); unknown keys are ignored like the reference."""
    overrides: dict[str, object] = {}
    typed = None
    return_audio = False
    if settings is not None:
        for key, value in settings.items():
            if key in _BOOL_KEYS:
                overrides[key] = _extract_bool(key, value)
            elif key in _F64_KEYS:
                overrides[key] = _extract_f64(key, value)
            elif key == "return_output_audio":
                return_audio = _extract_bool(key, value)
            elif key == "eq_bands_v2":
                typed = value
            elif key == "input_stage":  # new optional key (absent = reference behaviour)
                overrides[key] = value
    if typed is not None:
        overrides["use_typed_bands"] = True
    return abi.make_settings(**overrides), typed, return_audio


def _legacy_bands(bands: Sequence[tuple[float, float, float]]):
    bands = list(bands)
    if len(bands) != NUM_BANDS:
        raise ValueError(f"expected {NUM_BANDS} EQ bands, got {len(bands)}")
    return abi.legacy_bands(bands)


def _typed_bands(bands):
    bands = list(bands)
    if len(bands) != NUM_BANDS:
        raise ValueError(f"expected {NUM_BANDS} EQ bands, got {len(bands)}")
    for index, band in enumerate(bands):
        if band[0] not in abi.FILTER_IDS:  # lib.rs:170-175
            raise ValueError(f"band {index} has unsupported EQ filter type: {band[0]}")
    return abi.typed_bands(bands)


def _audio_1d(audio) -> np.ndarray:
    arr = np.asarray(audio)
    if arr.dtype != np.float32 or arr.ndim != 1:
        raise TypeError("audio must be a one-dimensional float32 array")
    if not arr.flags.c_contiguous:  # PyReadonlyArray1::as_slice (python_api.rs:505)
        raise ValueError("audio must be C-contiguous")
    return arr


def _check_rate_chain(sample_rate: float) -> None:
    if not np.isfinite(sample_rate) or sample_rate <= 0.0:
        raise ValueError("sample_rate must be positive and finite")


# ---- the reference's pyfunctions ---------------------------------------------------------------------------

def simulate_auto_eq_chain(audio, sample_rate: float, bands, settings: Mapping[str, object] | None = None,
                           *, device: int = 0) -> dict[str, Any]:
    """python_api.rs:378-714."""
    arr = _audio_1d(audio)
    sample_rate = float(sample_rate)
    _check_rate_chain(sample_rate)
    legacy = _legacy_bands(bands)
    st, typed, return_audio = settings_from_mapping(settings)
    band_arr = _typed_bands(typed) if typed is not None else legacy
    m, out = simulator(device).chain_render(arr, sample_rate, band_arr, st, return_audio=return_audio)
    result = abi.metrics_to_dict(m)
    if return_audio:
        result["output_audio"] = out.tolist()
    return result


def simulate_auto_eq_chain_batch(passages, sample_rate: float, candidates, *, pair_passage=None, pair_candidate=None,
                                 return_output_audio: bool = False, device: int = 0) -> list[dict[str, Any]]:
    """Batched form: ``candidates`` = sequence of ``(bands, settings)`` as simulate_auto_eq_chain takes them.

    Default pairing is the full cross product, candidate-major (pair = candidate * n_passages + passage).
    What apply_headroom_validation (headroom.py:306-320) and _calibrate_compressor_threshold
    (voice_setup.py:916-966) evaluate one call at a time.  With ``return_output_audio`` every dict carries
    ``"output_audio"`` (python_api.rs:701-712) -- as a float32 array, not a list: the callers
    (voice_setup.py:1535-1536) pass it to ``np.asarray(..., dtype=np.float32)`` either way."""
    sample_rate = float(sample_rate)
    _check_rate_chain(sample_rate)
    arrs = [_audio_1d(p) for p in passages]
    cands = (abi.AfCandidate * max(len(candidates), 1))()
    for i, (bands, settings) in enumerate(candidates):
        legacy = _legacy_bands(bands)
        st, typed, _ = settings_from_mapping(settings)
        band_arr = _typed_bands(typed) if typed is not None else legacy
        for b in range(NUM_BANDS):
            cands[i].bands[b] = band_arr[b]
        cands[i].settings = st
    cands_view = (abi.AfCandidate * len(candidates)).from_buffer(cands) if len(candidates) else cands
    metrics, audio = simulator(device).chain_sweep(arrs, sample_rate, cands_view, pair_passage, pair_candidate,
                                                   return_audio=bool(return_output_audio))
    n = len(arrs) * len(candidates) if pair_passage is None else len(pair_passage)
    out = [abi.metrics_to_dict(metrics[i]) for i in range(n)]
    if return_output_audio:
        for i in range(n):
            out[i]["output_audio"] = audio[i]
    return out


_MAKEUP_F64 = ("threshold_db", "ratio", "attack_ms", "release_ms", "makeup_gain_db", "target_lufs", "vad_reliability")
_MAKEUP_BOOL = ("adaptive_release", "sidechain_highpass_enabled")


def simulate_auto_makeup_control(audio, sample_rate: float, vad_probabilities, noise_floor_db: float,
                                 noise_reliability: float, settings: Mapping[str, object] | None = None,
                                 *, device: int = 0) -> dict[str, Any]:
    """python_api.rs:118-276: the compressor's auto-makeup controller over 480-sample control blocks.

    The runtime percentiles of the reference (wall time of each CPU block) have no per-block counterpart in one
    GPU pass; the three keys report the call's wall time divided over the blocks."""
    import time

    sample_rate = float(sample_rate)
    if not np.isfinite(sample_rate) or sample_rate <= 0.0:
        raise ValueError("sample_rate must be positive and finite")
    noise_floor_db, noise_reliability = float(noise_floor_db), float(noise_reliability)
    if not np.isfinite(noise_floor_db) or not np.isfinite(noise_reliability) or not 0.0 <= noise_reliability <= 1.0:
        raise ValueError("noise evidence must be finite and reliability must be between 0 and 1")
    vad = [float(v) for v in vad_probabilities]
    if any((not np.isfinite(v)) or not 0.0 <= v <= 1.0 for v in vad):
        raise ValueError("VAD probabilities must be finite and between 0 and 1")
    arr = _audio_1d(audio)
    overrides: dict[str, object] = {}
    return_audio = False
    if settings is not None:
        for key, value in settings.items():
            if key in _MAKEUP_F64:
                overrides[key] = _extract_f64(key, value)
            elif key in _MAKEUP_BOOL:
                overrides[key] = _extract_bool(key, value)
            elif key == "return_output_audio":
                return_audio = _extract_bool(key, value)
    st = abi.make_makeup_settings(**overrides)
    started = time.perf_counter()
    traces, out = simulator(device).auto_makeup_control(arr, sample_rate, vad, noise_floor_db, noise_reliability, st,
                                                        return_audio=return_audio)
    elapsed_ms = (time.perf_counter() - started) * 1000.0
    blocks = traces.shape[1]
    per_block_ms = elapsed_ms / blocks if blocks else 0.0
    result: dict[str, Any] = {
        "control_block_size": abi.MAKEUP_CONTROL_BLOCK,
        "control_cadence_hz": sample_rate / abi.MAKEUP_CONTROL_BLOCK,
        "processed_samples": int(arr.size),
    }
    for name, row in zip(abi.MAKEUP_TRACES, traces):
        result[name] = [float(v) for v in row]
    # The reference times every 480-sample block on its own (python_api.rs:224-226,257-271).  Here all blocks of the
    # capture are rendered by stage kernels that each cover many blocks: a per-block time does not exist, so the three
    # keys carry the AMORTISED cost per block (wall time of the call / blocks) and the flag below says so.
    result["p95_block_runtime_ms"] = per_block_ms
    result["p99_block_runtime_ms"] = per_block_ms
    result["max_block_runtime_ms"] = per_block_ms
    result["block_runtime_is_amortized"] = True
    result["runtime_ms"] = elapsed_ms
    if return_audio:
        result["output_audio"] = out.tolist()
    return result


def simulate_eq_v2(audio, sample_rate: float, bands, return_output_audio: bool = False, *, device: int = 0):
    """lib.rs:214-288."""
    sample_rate = float(sample_rate)
    if not np.isfinite(sample_rate) or sample_rate <= 0.0:
        raise ValueError("sample_rate must be finite and positive")
    band_arr = _typed_bands(bands)
    arr = _audio_1d(audio)
    st, out = simulator(device).eq_render(arr, sample_rate, band_arr, return_audio=return_output_audio)
    result: dict[str, Any] = {
        "input_sample_peak": float(st.input_sample_peak),
        "output_sample_peak": float(st.output_sample_peak),
        "input_true_peak": float(st.input_true_peak),
        "output_true_peak": float(st.output_true_peak),
        "input_rms": float(st.input_rms),
        "output_rms": float(st.output_rms),
        "max_response_db": float(st.max_response_db),
        "runtime_ms": float(st.runtime_ms),
        "sample_count": int(st.sample_count),
        "algorithmic_latency_samples": int(st.algorithmic_latency_samples),
        "non_finite_output": bool(st.non_finite_output),
    }
    if return_output_audio:
        result["output_audio"] = out.tolist()
    return result


def eq_magnitude_response(frequencies_hz, bands, sample_rate: float, *, device: int = 0) -> list[float]:
    """lib.rs:99-150 (legacy 3-tuple bands)."""
    sample_rate = float(sample_rate)
    if not np.isfinite(sample_rate) or sample_rate <= 0.0:
        raise ValueError("sample_rate must be finite and positive")
    band_arr = _legacy_bands(bands)
    return simulator(device).eq_response(list(frequencies_hz), band_arr, sample_rate, typed=False)[0].tolist()


def eq_magnitude_response_v2(frequencies_hz, bands, sample_rate: float, *, device: int = 0) -> list[float]:
    """lib.rs:191-212 (typed 6-tuple bands)."""
    sample_rate = float(sample_rate)
    if not np.isfinite(sample_rate) or sample_rate <= 0.0:
        raise ValueError("sample_rate must be finite and positive")
    band_arr = _typed_bands(bands)
    return simulator(device).eq_response(list(frequencies_hz), band_arr, sample_rate, typed=True)[0].tolist()


# ---- product resampler simulator (rust-core/src/audio/processor/resampling.rs:170-272) -------------------------
def _resampler_spec(input_rate, output_rate, chunk_size, sinc_len, window) -> abi.AfResamplerSpec:
    if window is not None and window not in abi.RESAMPLER_WINDOWS:
        # the reference validates rates, chunk and sinc_len first (resampling.rs:187-202); let the library do that
        # with a window id it rejects, then report the name as the reference does
        spec = native.resampler_spec(int(input_rate), int(output_rate), int(chunk_size), sinc_len, None)
        native.resampler_shape(spec, 0)
        raise ValueError(f"unsupported resampler window {str(window)!r}".replace("'", '"'))
    return native.resampler_spec(int(input_rate), int(output_rate), int(chunk_size), sinc_len, window)


def simulate_product_resampler_batch(signals, input_rate: int, output_rate: int, chunk_size: int = 1024,
                                     sinc_len: int | None = None, window: str | None = None, *, device: int = 0):
    """Batched form: `signals` [n_streams, n_in] (equal lengths) through the product resampler in ONE call ->
    (outputs [n_streams, frames] float64 ndarray, output_delay, expected_frames)."""
    spec = _resampler_spec(input_rate, output_rate, chunk_size, sinc_len, window)
    arr = np.asarray(signals, dtype=np.float64)
    if arr.ndim != 2:
        raise ValueError("signals must be a 2-d array [n_streams, n_in]")
    try:
        out, shape = simulator(device).product_resampler(arr, spec)
    except ValueError as error:
        if str(error).startswith("resampler flush"):
            raise RuntimeError(str(error)) from None  # PyRuntimeError in the reference (resampling.rs:251-255)
        raise
    return out, int(shape.delay), int(shape.expected_frames)


def simulate_product_resampler(samples, input_rate: int, output_rate: int, chunk_size: int = 1024,
                               sinc_len: int | None = None, window: str | None = None, *, device: int = 0):
    """resampling.rs:170-262 -> (output, output_delay, expected_frames, block_times_ns).  `block_times_ns` (the
    reference's wall clock per rubato process call) has no per-block counterpart on the GPU: every entry carries the
    amortised time of the whole call, one entry per block the reference would have processed."""
    import time
    arr = np.asarray(samples, dtype=np.float64).reshape(1, -1)
    started = time.perf_counter_ns()
    out, delay, expected = simulate_product_resampler_batch(arr, input_rate, output_rate, chunk_size, sinc_len, window,
                                                            device=device)
    elapsed = time.perf_counter_ns() - started
    spec = _resampler_spec(input_rate, output_rate, chunk_size, sinc_len, window)
    blocks = int(native.resampler_shape(spec, arr.shape[1]).blocks)
    return out[0].tolist(), delay, expected, [max(1, elapsed // max(blocks, 1))] * blocks


def product_resampler_configuration():
    """resampling.rs:263-272: (sinc_len, window, interpolation, oversampling_factor, chunk_size)."""
    return (128, "blackman", "cubic", 256, 1024)


# ---- names mic_eq/__init__.py:55-58 reads unguarded; the live engine is out of scope here -------------------
def _out_of_scope(*_args, **_kwargs):
    raise ImportError("the live audio engine is not part of the B200 chain-simulator build")


class AudioProcessor:  # noqa: D101
    def __init__(self, *args, **kwargs):
        _out_of_scope()


class DeviceInfo:  # noqa: D101
    def __init__(self, *args, **kwargs):
        _out_of_scope()


def list_input_devices():
    return []


def list_output_devices():
    return []
