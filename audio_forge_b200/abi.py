"""ctypes mirrors of the POD structs in ``include/afsim.h``.

These are the boundary definitions only (layout must match the C header field for
field); no DSP lives here.  Field names follow the reference's flat settings keys
(rust-core/src/audio/processor/python_api.rs:415-487) and its result dict
(python_api.rs:649-713).
"""
from __future__ import annotations

import ctypes as C

NUM_BANDS = 10

AFSIM_OK = 0
AFSIM_INVALID_ARGUMENT = 1
AFSIM_CUDA_ERROR = 2
AFSIM_OUT_OF_MEMORY = 3
AFSIM_UNSUPPORTED = 4

# Stable filter ids / names, reference rust-core/src/dsp/eq.rs:46-80
FILTER_NAMES = ("low_shelf", "bell", "high_shelf", "notch", "high_pass", "low_pass")
FILTER_IDS = {name: index for index, name in enumerate(FILTER_NAMES)}

INPUT_STAGES = {"none": 0, "dc_hp80": 1, "off": 1, "gentle": 2, "strong": 3}

DEFAULT_FREQUENCIES = (80.0, 160.0, 320.0, 640.0, 1280.0, 2500.0, 5000.0, 8000.0, 12000.0, 16000.0)
DEFAULT_Q = 1.41


class AfBand(C.Structure):
    _fields_ = [
        ("frequency_hz", C.c_double),
        ("gain_db", C.c_double),
        ("q", C.c_double),
        ("filter_type", C.c_uint8),
        ("slope_db_per_octave", C.c_uint8),
        ("enabled", C.c_uint8),
        ("reserved", C.c_uint8 * 5),
    ]


_BOOL_KEYS = (
    "use_typed_bands",
    "eq_before_deesser",
    "deesser_enabled",
    "deesser_auto_enabled",
    "compressor_enabled",
    "compressor_adaptive_release",
    "compressor_auto_makeup_enabled",
    "compressor_sidechain_highpass_enabled",
    "limiter_enabled",
    "limiter_careful_output_enabled",
)
_F64_KEYS = (
    "deesser_auto_amount",
    "deesser_low_cut_hz",
    "deesser_high_cut_hz",
    "deesser_threshold_db",
    "deesser_ratio",
    "deesser_attack_ms",
    "deesser_release_ms",
    "deesser_max_reduction_db",
    "compressor_threshold_db",
    "compressor_ratio",
    "compressor_attack_ms",
    "compressor_release_ms",
    "compressor_makeup_gain_db",
    "compressor_base_release_ms",
    "compressor_target_lufs",
    "limiter_ceiling_db",
    "limiter_release_ms",
    "limiter_lookahead_ms",
)


class AfChainSettings(C.Structure):
    _fields_ = (
        [(key, C.c_uint8) for key in _BOOL_KEYS]
        + [("input_stage", C.c_uint8), ("reserved", C.c_uint8 * 5)]
        + [(key, C.c_double) for key in _F64_KEYS]
    )


class AfCandidate(C.Structure):
    _fields_ = [("bands", AfBand * NUM_BANDS), ("settings", AfChainSettings)]


_METRIC_F32 = (
    "input_sample_peak_db",
    "input_rms_db",
    "output_sample_peak_db",
    "pre_limiter_true_peak_db",
    "output_true_peak_db",
    "output_rms_db",
    "limiter_effective_ceiling_db",
    "sample_headroom_db",
    "pre_limiter_true_peak_headroom_db",
    "true_peak_headroom_db",
    "limiter_gain_reduction_db",
    "true_peak_limiter_gain_reduction_db",
    "compressor_gain_reduction_db",
    "deesser_gain_reduction_db",
    "compressor_gain_reduction_median_db",
    "compressor_gain_reduction_p95_db",
    "compressor_gain_reduction_active_ratio",
    "active_output_gain_db",
    "silence_output_gain_db",
    "silence_level_delta_db",
    "compressor_pumping_score_db",
    "deesser_gain_reduction_median_db",
    "deesser_gain_reduction_p95_db",
    "analysis_block_ms",
    "active_analysis_threshold_db",
)


class AfChainMetrics(C.Structure):
    _fields_ = (
        [(key, C.c_float) for key in _METRIC_F32]
        + [
            ("non_finite_output", C.c_uint32),
            ("true_peak_limited_events", C.c_uint64),
            ("active_analysis_block_count", C.c_uint64),
            ("processed_samples", C.c_uint64),
            ("candidate_runtime_ms", C.c_double),
        ]
    )


class AfAutoMakeupSettings(C.Structure):
    """settings dict of simulate_auto_makeup_control (python_api.rs:168-192)."""
    _fields_ = [
        ("threshold_db", C.c_double),
        ("ratio", C.c_double),
        ("attack_ms", C.c_double),
        ("release_ms", C.c_double),
        ("makeup_gain_db", C.c_double),
        ("target_lufs", C.c_double),
        ("vad_reliability", C.c_double),
        ("adaptive_release", C.c_uint8),
        ("sidechain_highpass_enabled", C.c_uint8),
        ("reserved", C.c_uint8 * 6),
    ]


MAKEUP_CONTROL_BLOCK = 480
MAKEUP_TRACES = ("makeup_gain_db", "activity", "reliability", "gain_reduction_db", "input_rms_db", "output_rms_db")


def make_makeup_settings(**overrides) -> AfAutoMakeupSettings:
    """Reference defaults (python_api.rs:168-192) with overrides."""
    values = dict(threshold_db=-24.0, ratio=3.0, attack_ms=10.0, release_ms=180.0, makeup_gain_db=0.0, target_lufs=-18.0,
                  vad_reliability=1.0, adaptive_release=True, sidechain_highpass_enabled=True)
    unknown = set(overrides) - set(values)
    if unknown:
        raise KeyError(f"unknown auto-makeup settings: {sorted(unknown)}")
    values.update(overrides)
    s = AfAutoMakeupSettings()
    for key, value in values.items():
        setattr(s, key, int(bool(value)) if key in ("adaptive_release", "sidechain_highpass_enabled") else float(value))
    return s


class AfEqRenderStats(C.Structure):
    _fields_ = [
        ("input_sample_peak", C.c_float),
        ("output_sample_peak", C.c_float),
        ("input_true_peak", C.c_float),
        ("output_true_peak", C.c_float),
        ("input_rms", C.c_double),
        ("output_rms", C.c_double),
        ("max_response_db", C.c_double),
        ("runtime_ms", C.c_double),
        ("sample_count", C.c_uint64),
        ("algorithmic_latency_samples", C.c_uint64),
        ("non_finite_output", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


# Reference defaults of the flat settings dict (python_api.rs:415-487).
SETTINGS_DEFAULTS: dict[str, object] = {
    "use_typed_bands": False,
    "eq_before_deesser": False,
    "deesser_enabled": False,
    "deesser_auto_enabled": True,
    "compressor_enabled": True,
    "compressor_adaptive_release": False,
    "compressor_auto_makeup_enabled": False,
    "compressor_sidechain_highpass_enabled": True,
    "limiter_enabled": True,
    "limiter_careful_output_enabled": True,
    "input_stage": 0,
    "deesser_auto_amount": 0.5,
    "deesser_low_cut_hz": 4000.0,
    "deesser_high_cut_hz": 11000.0,
    "deesser_threshold_db": -28.0,
    "deesser_ratio": 4.0,
    "deesser_attack_ms": 2.0,
    "deesser_release_ms": 80.0,
    "deesser_max_reduction_db": 6.0,
    "compressor_threshold_db": -20.0,
    "compressor_ratio": 4.0,
    "compressor_attack_ms": 10.0,
    "compressor_release_ms": 200.0,
    "compressor_makeup_gain_db": 0.0,
    "compressor_base_release_ms": 50.0,
    "compressor_target_lufs": -18.0,
    "limiter_ceiling_db": -0.5,
    "limiter_release_ms": 50.0,
    "limiter_lookahead_ms": 2.0,
}


def make_settings(**overrides: object) -> AfChainSettings:
    """Build an AfChainSettings from the reference defaults plus overrides."""
    values = dict(SETTINGS_DEFAULTS)
    unknown = set(overrides) - set(values)
    if unknown:
        raise KeyError(f"unknown chain settings: {sorted(unknown)}")
    values.update(overrides)
    s = AfChainSettings()
    for key in _BOOL_KEYS:
        setattr(s, key, 1 if values[key] else 0)
    stage = values["input_stage"]
    s.input_stage = INPUT_STAGES[stage] if isinstance(stage, str) else int(stage)
    for key in _F64_KEYS:
        setattr(s, key, float(values[key]))  # type: ignore[arg-type]
    return s


def default_bands() -> "C.Array[AfBand]":
    """The reference's default 10-band layout (dsp/eq.rs:11-23,127-140)."""
    arr = (AfBand * NUM_BANDS)()
    for i in range(NUM_BANDS):
        arr[i].frequency_hz = DEFAULT_FREQUENCIES[i]
        arr[i].gain_db = 0.0
        arr[i].q = DEFAULT_Q
        arr[i].filter_type = 0 if i == 0 else 2 if i == NUM_BANDS - 1 else 1
        arr[i].slope_db_per_octave = 12
        arr[i].enabled = 1
    return arr


def legacy_bands(bands) -> "C.Array[AfBand]":
    """10 ``(frequency, gain_db, q)`` tuples -> AfBand[10] (legacy 3-tuple path)."""
    bands = list(bands)
    arr = default_bands()
    for i, (f, g, q) in enumerate(bands[:NUM_BANDS]):
        arr[i].frequency_hz = float(f)
        arr[i].gain_db = float(g)
        arr[i].q = float(q)
    return arr


def typed_bands(bands) -> "C.Array[AfBand]":
    """10 ``(name, f, g, q, slope, enabled)`` tuples -> AfBand[10]; unknown names map to id 255."""
    arr = (AfBand * NUM_BANDS)()
    for i, (name, f, g, q, slope, enabled) in enumerate(list(bands)[:NUM_BANDS]):
        arr[i].frequency_hz = float(f)
        arr[i].gain_db = float(g)
        arr[i].q = float(q)
        arr[i].filter_type = FILTER_IDS.get(name, 255)
        arr[i].slope_db_per_octave = int(slope)
        arr[i].enabled = 1 if enabled else 0
    return arr


def metrics_to_dict(m: AfChainMetrics) -> dict[str, object]:
    """AfChainMetrics -> the dict simulate_auto_eq_chain returns (python_api.rs:649-713)."""
    out: dict[str, object] = {key: float(getattr(m, key)) for key in _METRIC_F32}
    out["true_peak_limited_events"] = int(m.true_peak_limited_events)
    out["non_finite_output"] = bool(m.non_finite_output)
    out["candidate_runtime_ms"] = float(m.candidate_runtime_ms)
    out["active_analysis_block_count"] = int(m.active_analysis_block_count)
    out["processed_samples"] = int(m.processed_samples)
    return out


METRIC_F32_KEYS = _METRIC_F32


# sizeof(afsim::CandidateParams) (csrc/afsim_params.h): what one planned candidate costs on the H2D path
CANDIDATE_PARAMS_BYTES = 24 + 8 * (40 * 5 + 52 + 30 + 15 + 15 + 13 + 2 + 5)


def ctypes_sizeof_metrics() -> int:
    return C.sizeof(AfChainMetrics)


def ctypes_sizeof_candidate_params() -> int:
    return CANDIDATE_PARAMS_BYTES


# ---- product resampler simulator (include/afsim.h; rust-core/src/audio/processor/resampling.rs:158-272) ----
RESAMPLER_WINDOWS = ("blackman_harris", "blackman_harris_squared", "blackman", "blackman_squared", "hann", "hann_squared")


class AfResamplerSpec(C.Structure):
    _fields_ = [
        ("input_rate", C.c_uint32),
        ("output_rate", C.c_uint32),
        ("chunk_size", C.c_uint32),
        ("sinc_len", C.c_uint32),
        ("window", C.c_int32),
    ]


class AfResamplerShape(C.Structure):
    _fields_ = [
        ("frames", C.c_uint64),
        ("expected_frames", C.c_uint64),
        ("delay", C.c_uint32),
        ("blocks", C.c_uint32),
    ]
