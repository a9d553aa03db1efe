"""Test harness: the product's stage-kernel bodies compiled for the CPU (see hostsim.cpp).

TEST INFRASTRUCTURE ONLY.  Lets the CPU test-suite check the chunk x stage schedule, state
parking, ring addressing and the finalize reduction against the oracle without a GPU.  The
product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from audio_forge_b200 import abi

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "libhostsim.so"
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", str(_HERE), "libhostsim.so"], check=True, capture_output=True)
        L = C.CDLL(str(_LIB))
        f32p = C.POINTER(C.c_float)
        u32p = C.POINTER(C.c_uint32)
        L.hostsim_last_error.restype = C.c_char_p
        L.hostsim_chain_sweep.argtypes = [C.POINTER(f32p), C.POINTER(C.c_size_t), C.c_size_t, C.c_double,
                                          C.POINTER(abi.AfCandidate), C.c_size_t, u32p, u32p, C.c_size_t, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.POINTER(abi.AfChainMetrics), f32p, f32p]
        L.hostsim_makeup_control.argtypes = [f32p, C.c_size_t, C.c_double, C.POINTER(C.c_double), C.c_size_t, C.c_double,
                                             C.c_double, C.POINTER(abi.AfAutoMakeupSettings), C.c_int, C.c_int, C.c_int,
                                             f32p, f32p]
        L.hostsim_cut_group.argtypes = [u32p, u32p, C.c_size_t, C.c_int, u32p, u32p]
        L.hostsim_eq_scan.argtypes = [f32p, C.c_size_t, C.c_double, C.POINTER(abi.AfBand), C.c_int, f32p]
        _lib = L
    return _lib


class HostsimError(ValueError):
    pass


def chain_sweep(passages, sample_rate, candidates, pair_passage, pair_candidate, *, chunk=1024, slots=2, eq_k=5, split=0,
                want_audio=False, want_rows=False):
    """-> (AfChainMetrics array, audio [n_pairs, T] | None, rows [4, n_rows, n_pairs] | None)."""
    passages = [np.ascontiguousarray(p, dtype=np.float32) for p in passages]
    f32p = C.POINTER(C.c_float)
    ptrs = (f32p * len(passages))(*[p.ctypes.data_as(f32p) for p in passages])
    lens = (C.c_size_t * len(passages))(*[p.size for p in passages])
    pp = np.ascontiguousarray(pair_passage, dtype=np.uint32)
    pc = np.ascontiguousarray(pair_candidate, dtype=np.uint32)
    n_pairs = pp.size
    T = passages[int(pp[0])].size if n_pairs else 0
    out = (abi.AfChainMetrics * max(n_pairs, 1))()
    audio = np.zeros((n_pairs, T), dtype=np.float32) if want_audio else None
    block = max(1, min(8192, int(round(sample_rate * 0.020))))
    n_rows = (T + block - 1) // block
    rows = np.zeros((4, n_rows, n_pairs), dtype=np.float32) if want_rows else None
    rc = lib().hostsim_chain_sweep(ptrs, lens, len(passages), float(sample_rate), candidates, len(candidates),
                                   pp.ctypes.data_as(C.POINTER(C.c_uint32)), pc.ctypes.data_as(C.POINTER(C.c_uint32)),
                                   n_pairs, int(chunk), int(slots), int(eq_k), int(split), out,
                                   audio.ctypes.data_as(f32p) if audio is not None else None,
                                   rows.ctypes.data_as(f32p) if rows is not None else None)
    if rc != abi.AFSIM_OK:
        raise HostsimError(lib().hostsim_last_error().decode())
    return out, audio, rows


def makeup_control(audio, sample_rate, vad, noise_floor_db, noise_reliability, settings, *, chunk=960, slots=2,
                   direct=False, want_audio=False):
    """simulate_auto_makeup_control through the product's stage bodies -> (traces [6, blocks], audio | None)."""
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    vad = np.ascontiguousarray(vad if vad is not None else [], dtype=np.float64)
    blocks = (audio.size + 479) // 480
    traces = np.zeros((6, blocks), dtype=np.float32)
    out = np.zeros_like(audio) if want_audio else None
    f32p = C.POINTER(C.c_float)
    rc = lib().hostsim_makeup_control(audio.ctypes.data_as(f32p), audio.size, float(sample_rate),
                                      vad.ctypes.data_as(C.POINTER(C.c_double)) if vad.size else None, vad.size,
                                      float(noise_floor_db), float(noise_reliability), C.byref(settings), int(chunk),
                                      int(slots), 1 if direct else 0, traces.ctypes.data_as(f32p),
                                      out.ctypes.data_as(f32p) if out is not None else None)
    if rc != abi.AFSIM_OK:
        raise HostsimError(lib().hostsim_last_error().decode())
    return traces, out


def eq_scan(audio, sample_rate, bands, log2_len=6):
    """The time-parallel EQ render (afsim_eqscan.h) walked on the host -> audio."""
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    out = np.zeros_like(audio)
    f32p = C.POINTER(C.c_float)
    rc = lib().hostsim_eq_scan(audio.ctypes.data_as(f32p), audio.size, float(sample_rate), bands, int(log2_len),
                               out.ctypes.data_as(f32p))
    if rc != abi.AFSIM_OK:
        raise HostsimError(lib().hostsim_last_error().decode())
    return out


def cut_group(passage, eq_class, max_streams):
    """-> (piece index, position inside the piece) of every stream, number of pieces (cut_stream_group, afsim_plan.cpp)."""
    p = np.ascontiguousarray(passage, dtype=np.uint32)
    c = np.ascontiguousarray(eq_class, dtype=np.uint32)
    piece = np.full(p.size, 0xFFFFFFFF, dtype=np.uint32)
    pos = np.full(p.size, 0xFFFFFFFF, dtype=np.uint32)
    u32p = C.POINTER(C.c_uint32)
    n = lib().hostsim_cut_group(p.ctypes.data_as(u32p), c.ctypes.data_as(u32p), p.size, int(max_streams),
                                piece.ctypes.data_as(u32p), pos.ctypes.data_as(u32p))
    return piece, pos, n
