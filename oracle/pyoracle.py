"""ctypes loader for the parity oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the product
package (audio_forge_b200).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from audio_forge_b200 import abi

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"


def build(force: bool = False) -> Path:
    """Compile liboracle.so with the committed Makefile (g++ only)."""
    sources = [_HERE / "afsim_oracle_capi.cpp", _HERE / "afsim_oracle.hpp", _HERE / "true_peak_fir.inc"]
    stale = force or not _LIB_PATH.exists() or any(
        s.stat().st_mtime > _LIB_PATH.stat().st_mtime for s in sources
    )
    if stale:
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        f32p = C.POINTER(C.c_float)
        f64p = C.POINTER(C.c_double)
        L.orc_last_error.restype = C.c_char_p
        L.orc_chain_render.argtypes = [f32p, C.c_size_t, C.c_double, C.POINTER(abi.AfBand),
                                       C.POINTER(abi.AfChainSettings), C.POINTER(abi.AfChainMetrics), f32p, f32p,
                                       C.c_size_t]
        L.orc_chain_sweep.argtypes = [C.POINTER(f32p), C.POINTER(C.c_size_t), C.c_double, C.POINTER(abi.AfCandidate),
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_size_t,
                                      C.POINTER(abi.AfChainMetrics), C.c_int]
        L.orc_eq_render.argtypes = [f32p, C.c_size_t, C.c_double, C.POINTER(abi.AfBand),
                                    C.POINTER(abi.AfEqRenderStats), f32p]
        L.orc_eq_response.argtypes = [f64p, C.c_size_t, C.POINTER(abi.AfBand), C.c_int, C.c_double, f64p]
        L.orc_auto_makeup_control.argtypes = [f32p, C.c_size_t, C.c_double, f64p, C.c_size_t, C.c_double, C.c_double,
                                              C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                              C.c_int, C.c_int, C.c_double, f32p, f32p]
        # object-level API
        L.orc_proc_new.restype = C.c_void_p
        L.orc_proc_new.argtypes = [C.c_double]
        L.orc_proc_free.argtypes = [C.c_void_p]
        for name in ("deesser_enabled", "eq_enabled", "compressor_enabled", "limiter_enabled", "eq_before_deesser"):
            getattr(L, f"orc_proc_set_{name}").argtypes = [C.c_void_p, C.c_int]
        L.orc_proc_deesser_set_auto_enabled.argtypes = [C.c_void_p, C.c_int]
        L.orc_proc_comp_set_adaptive_release.argtypes = [C.c_void_p, C.c_int]
        for name in ("deesser_set_auto_amount", "deesser_set_max_reduction_db", "comp_set_threshold", "comp_set_ratio",
                     "comp_set_attack_time", "comp_set_release_time", "comp_set_makeup_gain", "limiter_set_ceiling",
                     "limiter_set_release_time"):
            getattr(L, f"orc_proc_{name}").argtypes = [C.c_void_p, C.c_double]
        for name in ("frequency", "gain", "q"):
            getattr(L, f"orc_proc_eq_set_band_{name}").argtypes = [C.c_void_p, C.c_size_t, C.c_double]
        L.orc_proc_process_block.argtypes = [C.c_void_p, f32p, C.c_size_t, f32p]
        L.orc_biquad_new.restype = C.c_void_p
        L.orc_biquad_new.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_biquad_free.argtypes = [C.c_void_p]
        L.orc_biquad_process.argtypes = [C.c_void_p, f32p, C.c_size_t]
        L.orc_biquad_response_db.restype = C.c_double
        L.orc_biquad_response_db.argtypes = [C.c_void_p, C.c_double]
        L.orc_biquad_set_gain_db.argtypes = [C.c_void_p, C.c_double]
        L.orc_biquad_set_frequency.argtypes = [C.c_void_p, C.c_double]
        L.orc_biquad_reset.argtypes = [C.c_void_p]
        L.orc_biquad_is_crossfading.argtypes = [C.c_void_p]
        L.orc_biquad_coeffs.argtypes = [C.c_void_p, f64p]
        L.orc_comp_new.restype = C.c_void_p
        L.orc_comp_new.argtypes = [C.c_double] * 7
        L.orc_comp_free.argtypes = [C.c_void_p]
        L.orc_comp_compute_gain_reduction.restype = C.c_double
        L.orc_comp_compute_gain_reduction.argtypes = [C.c_void_p, C.c_double]
        L.orc_comp_blended_detector_db.restype = C.c_double
        L.orc_comp_blended_detector_db.argtypes = [C.c_double, C.c_double]
        for name in ("set_adaptive_release", "set_sidechain_highpass_enabled", "set_auto_makeup_enabled"):
            getattr(L, f"orc_comp_{name}").argtypes = [C.c_void_p, C.c_int]
        L.orc_comp_process_block.argtypes = [C.c_void_p, f32p, C.c_size_t]
        L.orc_comp_process_samples.argtypes = [C.c_void_p, f32p, C.c_size_t]
        L.orc_comp_set_target_lufs.argtypes = [C.c_void_p, C.c_double]
        L.orc_comp_set_noise_reference_reliability.argtypes = [C.c_void_p, C.c_double]
        L.orc_comp_process_block_with_activity.argtypes = [C.c_void_p, f32p, C.c_size_t, C.c_int, C.c_double, C.c_double,
                                                           C.c_double, C.c_double]
        L.orc_meter_new.restype = C.c_void_p
        L.orc_meter_new.argtypes = [C.c_uint32]
        L.orc_meter_free.argtypes = [C.c_void_p]
        L.orc_meter_process.argtypes = [C.c_void_p, f32p, C.c_size_t]
        L.orc_meter_momentary.restype = C.c_float
        L.orc_meter_momentary.argtypes = [C.c_void_p]
        L.orc_meter_reset.argtypes = [C.c_void_p]
        L.orc_comp_update_auto_makeup_gain.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_size_t]
        L.orc_comp_estimate_activity.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, f64p]
        L.orc_comp_speech_activity_from_rms_db.restype = C.c_double
        L.orc_comp_speech_activity_from_rms_db.argtypes = [C.c_double]
        L.orc_comp_set_limiter_feedback_gain_reduction_db.argtypes = [C.c_void_p, C.c_double]
        L.orc_comp_set_makeup_gain.argtypes = [C.c_void_p, C.c_double]
        for name in ("auto_makeup_activity", "auto_makeup_activity_reliability"):
            getattr(L, f"orc_comp_{name}").restype = C.c_double
            getattr(L, f"orc_comp_{name}").argtypes = [C.c_void_p]
        for name in ("gain_reduction", "makeup_gain", "plosive_ratio"):
            getattr(L, f"orc_comp_{name}").restype = C.c_double
            getattr(L, f"orc_comp_{name}").argtypes = [C.c_void_p]
        L.orc_limiter_new.restype = C.c_void_p
        L.orc_limiter_new.argtypes = [C.c_double] * 4
        L.orc_limiter_free.argtypes = [C.c_void_p]
        L.orc_limiter_lookahead_samples.restype = C.c_size_t
        L.orc_limiter_lookahead_samples.argtypes = [C.c_void_p]
        L.orc_limiter_set_lookahead_ms.argtypes = [C.c_void_p, C.c_double]
        L.orc_limiter_process.argtypes = [C.c_void_p, f32p, C.c_size_t]
        L.orc_limiter_peak_gr_and_reset.restype = C.c_double
        L.orc_limiter_peak_gr_and_reset.argtypes = [C.c_void_p]
        L.orc_tpd_new.restype = C.c_void_p
        L.orc_tpd_free.argtypes = [C.c_void_p]
        L.orc_tpd_process.restype = C.c_float
        L.orc_tpd_process.argtypes = [C.c_void_p, f32p, C.c_size_t]
        L.orc_tpl_new.restype = C.c_void_p
        L.orc_tpl_new.argtypes = [C.c_float] * 3
        L.orc_tpl_free.argtypes = [C.c_void_p]
        L.orc_tpl_set_ceiling_linear.argtypes = [C.c_void_p, C.c_float]
        L.orc_tpl_process.argtypes = [C.c_void_p, f32p, C.c_size_t, f32p]
        L.orc_input_stage_process.argtypes = [C.c_int, C.c_double, f32p, C.c_size_t, f32p]
        L.orc_cleanup_harness.argtypes = [C.c_int, C.c_float, f32p, C.c_size_t, C.c_size_t, f32p]
        L.orc_cleanup_analyze.argtypes = [C.c_int, C.c_float, f32p, C.c_size_t, f32p]
        L.orc_percentile_f32.restype = C.c_float
        L.orc_percentile_f32.argtypes = [f32p, C.c_size_t, C.c_float]
        L.orc_pumping_score.restype = C.c_float
        L.orc_pumping_score.argtypes = [f32p, C.c_size_t, C.c_float]
        L.orc_time_constant_to_coeff.restype = C.c_double
        L.orc_time_constant_to_coeff.argtypes = [C.c_double, C.c_double]
        L.orc_butterworth_q.restype = C.c_double
        L.orc_butterworth_q.argtypes = [C.c_size_t, C.c_size_t]
        _lib = L
    return _lib


class OracleError(ValueError):
    pass


def fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _check(rc: int) -> None:
    if rc != abi.AFSIM_OK:
        raise OracleError(lib().orc_last_error().decode())


def chain_render(audio: np.ndarray, sample_rate: float, bands, settings: abi.AfChainSettings, *,
                 return_audio: bool = False, return_rows: bool = False):
    """simulate_auto_eq_chain restated on the CPU -> (AfChainMetrics, audio|None, rows|None)."""
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    m = abi.AfChainMetrics()
    out = np.zeros_like(audio) if return_audio else None
    block = max(1, min(8192, int(round(sample_rate * 0.020)))) if np.isfinite(sample_rate) and sample_rate > 0 else 1
    n_rows = (audio.size + block - 1) // block
    rows = np.zeros((n_rows, 4), dtype=np.float32) if return_rows else None
    rc = lib().orc_chain_render(fptr(audio), audio.size, float(sample_rate), bands, C.byref(settings), C.byref(m),
                                fptr(out) if out is not None else None,
                                fptr(rows) if rows is not None else None, n_rows)
    _check(rc)
    return m, out, rows


def chain_sweep(passages, sample_rate: float, candidates, pair_passage, pair_candidate, n_threads: int = 1):
    """Independent-stream CPU sweep (n_threads workers) -> array of AfChainMetrics."""
    passages = [np.ascontiguousarray(p, dtype=np.float32) for p in passages]
    f32p = C.POINTER(C.c_float)
    ptrs = (f32p * len(passages))(*[fptr(p) for p in passages])
    lens = (C.c_size_t * len(passages))(*[p.size for p in passages])
    pp = np.ascontiguousarray(pair_passage, dtype=np.uint32)
    pc = np.ascontiguousarray(pair_candidate, dtype=np.uint32)
    out = (abi.AfChainMetrics * pp.size)()
    rc = lib().orc_chain_sweep(ptrs, lens, float(sample_rate), candidates,
                               pp.ctypes.data_as(C.POINTER(C.c_uint32)), pc.ctypes.data_as(C.POINTER(C.c_uint32)),
                               pp.size, out, int(n_threads))
    _check(rc)
    return out


def eq_render(audio: np.ndarray, sample_rate: float, bands, *, return_audio: bool = False):
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    st = abi.AfEqRenderStats()
    out = np.zeros_like(audio) if return_audio else None
    rc = lib().orc_eq_render(fptr(audio), audio.size, float(sample_rate), bands, C.byref(st),
                             fptr(out) if out is not None else None)
    _check(rc)
    return st, out


def eq_response(frequencies_hz, bands, sample_rate: float, typed: bool) -> np.ndarray:
    freqs = np.ascontiguousarray(frequencies_hz, dtype=np.float64)
    out = np.zeros_like(freqs)
    rc = lib().orc_eq_response(dptr(freqs), freqs.size, bands, 1 if typed else 0, float(sample_rate), dptr(out))
    _check(rc)
    return out


def auto_makeup_control(audio, sample_rate: float, vad, noise_floor_db: float, noise_reliability: float,
                        settings: abi.AfAutoMakeupSettings, *, return_audio: bool = False):
    """simulate_auto_makeup_control (python_api.rs:118-276) -> (traces [6, blocks] float32, audio | None)."""
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    vad = np.ascontiguousarray(vad if vad is not None else [], dtype=np.float64)
    blocks = (audio.size + 479) // 480
    traces = np.zeros((6, blocks), dtype=np.float32)
    out = np.zeros_like(audio) if return_audio else None
    rc = lib().orc_auto_makeup_control(fptr(audio), audio.size, float(sample_rate), dptr(vad) if vad.size else None, vad.size,
                                       float(noise_floor_db), float(noise_reliability), settings.threshold_db, settings.ratio,
                                       settings.attack_ms, settings.release_ms, settings.makeup_gain_db, settings.target_lufs,
                                       int(settings.adaptive_release), int(settings.sidechain_highpass_enabled),
                                       settings.vad_reliability, fptr(traces), fptr(out) if out is not None else None)
    _check(rc)
    return traces, out
