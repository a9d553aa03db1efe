"""ctypes binding of libafsim.so (the sm_100a C-ABI library, include/afsim.h).

This is the only door from Python into the product.  There is no CPU fallback: when the
library is missing, or no sm_100 CUDA device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

from . import abi

# The chunk x stage wavefront runs every stage on its own CUDA stream; the default of 8 hardware work
# queues would serialise most of them.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libafsim.so"
_lib = None

EXPORTS = (
    "afsim_abi_version", "afsim_create", "afsim_destroy", "afsim_trim", "afsim_last_error", "afsim_create_error",
    "afsim_chain_settings_default", "afsim_default_bands", "afsim_chain_render", "afsim_eq_render",
    "afsim_eq_response", "afsim_chain_sweep", "afsim_sweep_prepare", "afsim_sweep_prepare_synthetic",
    "afsim_sweep_launch", "afsim_sweep_collect", "afsim_sweep_status", "afsim_sweep_collect_audio", "afsim_sweep_metrics_device_ptr",
    "afsim_sweep_kernel_count", "afsim_sweep_last_render_ms", "afsim_sweep_release",
    "afsim_sweep_profile_stages", "afsim_sweep_profile_wavefront", "afsim_sweep_batch_info", "afsim_measure_issue_peak",
    "afsim_multi_create", "afsim_multi_destroy", "afsim_multi_device_count", "afsim_multi_last_error",
    "afsim_multi_create_error", "afsim_multi_chain_sweep", "afsim_multi_partition",
    "afsim_selftest_math", "afsim_auto_makeup_settings_default", "afsim_auto_makeup_control", "afsim_auto_makeup_sweep",
    "afsim_resampler_spec_default", "afsim_product_resampler_shape", "afsim_product_resampler",
    "afsim_product_resampler_device", "afsim_product_resampler_plan",
)
STAGE_NAMES = ("input", "input_true_peak", "deesser", "eq", "compressor", "limiter", "output", "finalize",
               "comp_r1", "comp_m2", "comp_r3", "comp_m4", "comp_r5", "comp_m6", "lim_m", "lim_r", "tp_fir_in", "tp_r",
               "tp_fir_out", "de_ra", "de_mb", "de_rc", "comp_r7", "de_mc2", "de_rc3", "de_rc1a", "de_mc1b", "de_rc1c") + ("?",) * 12 + ("input_fanout", "tail")


class AfsimError(RuntimeError):
    """Library-level failure (CUDA error, out of memory, unsupported settings)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"afsim status {status}: {message}")
        self.status = status
        self.message = message


def build(force: bool = False) -> Path:
    """Compile libafsim.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    csrc = _PKG / "csrc"
    if force:
        subprocess.run(["make", "-C", str(csrc), "clean"], check=True, capture_output=True)
    proc = subprocess.run(["make", "-j4", "-C", str(csrc)], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libafsim.so failed:\n" + proc.stdout[-4000:] + proc.stderr[-4000:])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `make -C audio_forge_b200/csrc` (or __graft_entry__.build()); "
            "there is no CPU fallback")
    L = C.CDLL(str(LIB_PATH))
    f32p, f64p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    vp, szp = C.c_void_p, C.POINTER(C.c_size_t)
    bands_p, settings_p = C.POINTER(abi.AfBand), C.POINTER(abi.AfChainSettings)
    cand_p, metrics_p = C.POINTER(abi.AfCandidate), C.POINTER(abi.AfChainMetrics)
    L.afsim_abi_version.restype = C.c_int
    L.afsim_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    L.afsim_destroy.argtypes = [vp]
    L.afsim_destroy.restype = None
    L.afsim_trim.argtypes = [vp]
    L.afsim_last_error.argtypes = [vp]
    L.afsim_last_error.restype = C.c_char_p
    L.afsim_create_error.restype = C.c_char_p
    L.afsim_chain_settings_default.argtypes = [settings_p]
    L.afsim_chain_settings_default.restype = None
    L.afsim_default_bands.argtypes = [bands_p]
    L.afsim_default_bands.restype = None
    L.afsim_chain_render.argtypes = [vp, f32p, C.c_size_t, C.c_double, bands_p, settings_p, metrics_p, f32p]
    L.afsim_eq_render.argtypes = [vp, f32p, C.c_size_t, C.c_double, bands_p, C.POINTER(abi.AfEqRenderStats), f32p]
    L.afsim_eq_response.argtypes = [vp, f64p, C.c_size_t, bands_p, C.c_size_t, C.c_int, C.c_double, f64p]
    L.afsim_chain_sweep.argtypes = [vp, C.POINTER(f32p), szp, C.c_size_t, C.c_double, cand_p, C.c_size_t, u32p, u32p,
                                    C.c_size_t, metrics_p, C.POINTER(f32p)]
    L.afsim_sweep_prepare.argtypes = [vp, C.POINTER(f32p), szp, C.c_size_t, C.c_double, cand_p, C.c_size_t, u32p, u32p,
                                      C.c_size_t, C.c_int, C.POINTER(vp)]
    L.afsim_sweep_prepare_synthetic.argtypes = [vp, C.c_int, C.c_size_t, C.c_size_t, C.c_double, cand_p, C.c_size_t,
                                                u32p, u32p, C.c_size_t, C.c_int, C.POINTER(vp)]
    L.afsim_sweep_launch.argtypes = [vp, vp]
    L.afsim_sweep_collect.argtypes = [vp, vp, metrics_p]
    L.afsim_sweep_status.argtypes = [vp, vp]
    L.afsim_multi_create.argtypes = [C.c_uint32, C.POINTER(vp)]
    L.afsim_multi_destroy.argtypes = [vp]
    L.afsim_multi_destroy.restype = None
    L.afsim_multi_device_count.argtypes = [vp]
    L.afsim_multi_last_error.argtypes = [vp]
    L.afsim_multi_last_error.restype = C.c_char_p
    L.afsim_multi_create_error.restype = C.c_char_p
    L.afsim_multi_chain_sweep.argtypes = [vp, C.POINTER(f32p), szp, C.c_size_t, C.c_double, cand_p, C.c_size_t, u32p, u32p,
                                          C.c_size_t, metrics_p, f32p]
    L.afsim_multi_partition.argtypes = [cand_p, C.c_size_t, szp, C.c_size_t, u32p, u32p, C.c_size_t, C.c_int, u32p]
    L.afsim_sweep_collect_audio.argtypes = [vp, vp, C.c_size_t, f32p, C.c_size_t]
    L.afsim_sweep_metrics_device_ptr.argtypes = [vp]
    L.afsim_sweep_metrics_device_ptr.restype = vp
    L.afsim_sweep_kernel_count.argtypes = [vp]
    L.afsim_sweep_kernel_count.restype = C.c_int
    L.afsim_sweep_last_render_ms.argtypes = [vp, vp, f32p]
    L.afsim_sweep_release.argtypes = [vp, vp]
    L.afsim_sweep_release.restype = None
    L.afsim_sweep_profile_stages.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(C.c_int), f32p, C.POINTER(C.c_int),
                                             C.POINTER(C.c_int)]
    L.afsim_sweep_batch_info.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.afsim_measure_issue_peak.argtypes = [vp, C.c_int, f64p]
    L.afsim_sweep_profile_wavefront.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), f32p, f32p,
                                                C.POINTER(C.c_int)]
    L.afsim_selftest_math.argtypes = [vp, C.c_uint64, C.POINTER(C.c_uint64)]
    spec_p, shape_p = C.POINTER(abi.AfResamplerSpec), C.POINTER(abi.AfResamplerShape)
    L.afsim_resampler_spec_default.argtypes = [spec_p]
    L.afsim_resampler_spec_default.restype = None
    L.afsim_product_resampler_shape.argtypes = [spec_p, C.c_size_t, shape_p, C.c_char_p, C.c_size_t]
    L.afsim_product_resampler.argtypes = [vp, spec_p, C.POINTER(f64p), C.c_size_t, C.c_size_t, C.POINTER(f64p), shape_p]
    L.afsim_product_resampler_device.argtypes = [vp, spec_p, vp, C.c_size_t, C.c_size_t, C.c_size_t, vp, C.c_size_t, f32p]
    L.afsim_product_resampler_plan.argtypes = [spec_p, C.c_size_t, f64p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), f64p]
    mk_p = C.POINTER(abi.AfAutoMakeupSettings)
    L.afsim_auto_makeup_settings_default.argtypes = [mk_p]
    L.afsim_auto_makeup_settings_default.restype = None
    L.afsim_auto_makeup_control.argtypes = [vp, f32p, C.c_size_t, C.c_double, f64p, C.c_size_t, C.c_double, C.c_double,
                                            mk_p, f32p, f32p]
    L.afsim_auto_makeup_sweep.argtypes = [vp, C.POINTER(f32p), szp, C.c_size_t, C.c_double, C.POINTER(f64p), f64p, f64p,
                                          mk_p, C.POINTER(f32p), C.POINTER(f32p)]
    _lib = L
    return L


def _f32p(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32)) if a is not None else None


def resampler_spec(input_rate: int, output_rate: int, chunk_size: int = 1024, sinc_len: int | None = None,
                   window: str | None = None) -> abi.AfResamplerSpec:
    """AfResamplerSpec with the product defaults (128 taps, blackman: resampling.rs:131-138) for None."""
    spec = abi.AfResamplerSpec()
    lib().afsim_resampler_spec_default(C.byref(spec))
    if input_rate < 0 or output_rate < 0 or chunk_size < 0 or (sinc_len is not None and sinc_len < 0):
        raise OverflowError("can't convert negative int to unsigned")  # PyO3's u32 / usize extraction
    spec.input_rate, spec.output_rate, spec.chunk_size = int(input_rate), int(output_rate), min(int(chunk_size), 0xFFFFFFFF)
    if sinc_len is not None:
        spec.sinc_len = min(int(sinc_len), 0xFFFFFFFF)
    if window is not None:
        spec.window = abi.RESAMPLER_WINDOWS.index(window) if window in abi.RESAMPLER_WINDOWS else -1
    return spec


def resampler_shape(spec: abi.AfResamplerSpec, n_in: int, check: bool = True):
    """(frames, expected_frames, delay, blocks) of a render; host only.  check=False -> None instead of raising."""
    shape = abi.AfResamplerShape()
    err = C.create_string_buffer(256)
    rc = lib().afsim_product_resampler_shape(C.byref(spec), int(n_in), C.byref(shape), err, 256)
    if rc != abi.AFSIM_OK:
        if not check:
            return None
        if rc == abi.AFSIM_INVALID_ARGUMENT:
            raise ValueError(err.value.decode())
        raise AfsimError(rc, err.value.decode())
    return shape


def resampler_plan(spec: abi.AfResamplerSpec, n_in: int, with_table: bool = True):
    """The planner's phase table and frame list (audit hook of the CPU tests)."""
    shape = resampler_shape(spec, n_in)
    n = int(shape.frames)
    table = np.zeros((256, int(spec.sinc_len)), dtype=np.float64) if with_table else None
    base, phase, frac = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.float64)
    rc = lib().afsim_product_resampler_plan(C.byref(spec), int(n_in), table.ctypes.data_as(C.POINTER(C.c_double)) if with_table else None,
                                            base.ctypes.data_as(C.POINTER(C.c_int64)), phase.ctypes.data_as(C.POINTER(C.c_int32)),
                                            frac.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != abi.AFSIM_OK:
        raise AfsimError(rc, "afsim_product_resampler_plan")
    return shape, table, base, phase, frac


class Sweep:
    """A candidate x passage sweep resident in HBM (afsim_sweep_* in include/afsim.h)."""

    def __init__(self, sim: "Simulator", ptr: int, n_pairs: int, pair_len):
        self._sim, self._ptr, self.n_pairs, self._pair_len = sim, C.c_void_p(ptr), n_pairs, pair_len

    def launch(self) -> None:
        self._sim._check(lib().afsim_sweep_launch(self._sim._h, self._ptr))

    def collect(self):
        out = (abi.AfChainMetrics * max(self.n_pairs, 1))()
        self._sim._check(lib().afsim_sweep_collect(self._sim._h, self._ptr, out))
        return out

    def collect_audio(self, pair: int) -> np.ndarray:
        n = int(self._pair_len[pair])
        out = np.zeros(n, dtype=np.float32)
        self._sim._check(lib().afsim_sweep_collect_audio(self._sim._h, self._ptr, pair, _f32p(out), n))
        return out

    def render_ms(self) -> float:
        ms = C.c_float(0.0)
        self._sim._check(lib().afsim_sweep_last_render_ms(self._sim._h, self._ptr, C.byref(ms)))
        return float(ms.value)

    def profile_stages(self, max_chunks: int = 0):
        """Serialised pass with CUDA events around every stage launch -> [(stage name, total ms, launches)]."""
        cap = 32
        kinds, ms, launches, n = (C.c_int * cap)(), (C.c_float * cap)(), (C.c_int * cap)(), C.c_int(0)
        self._sim._check(lib().afsim_sweep_profile_stages(self._sim._h, self._ptr, int(max_chunks), cap, kinds, ms,
                                                          launches, C.byref(n)))
        return [(STAGE_NAMES[kinds[i]], float(ms[i]), int(launches[i])) for i in range(n.value)]

    def profile_wavefront(self, first_chunk: int = 64, n_chunks: int = 64):
        """Live wavefront with timing events -> [(stage name, busy ms per launch under contention, period ms)]."""
        cap = 32
        kinds, busy, period, n = (C.c_int * cap)(), (C.c_float * cap)(), (C.c_float * cap)(), C.c_int(0)
        self._sim._check(lib().afsim_sweep_profile_wavefront(self._sim._h, self._ptr, int(first_chunk), int(n_chunks), cap,
                                                             kinds, busy, period, C.byref(n)))
        return [(STAGE_NAMES[kinds[i]], float(busy[i]), float(period[i])) for i in range(n.value)]

    def batch_info(self) -> dict:
        """Shape of the first batch (what profile_stages / profile_wavefront time): streams per launch of every stage."""
        cap = 32
        info, streams, n = (C.c_int * 6)(), (C.c_int * cap)(), C.c_int(0)
        self._sim._check(lib().afsim_sweep_batch_info(self._ptr, cap, info, streams, C.byref(n)))
        return {"batches": info[0], "streams": info[1], "chunk": info[2], "slots": info[3], "stages": info[4],
                "samples": info[5], "stage_streams": [int(streams[i]) for i in range(n.value)]}

    @property
    def kernel_count(self) -> int:
        return int(lib().afsim_sweep_kernel_count(self._ptr))

    @property
    def metrics_device_ptr(self) -> int:
        return int(lib().afsim_sweep_metrics_device_ptr(self._ptr) or 0)

    def release(self) -> None:
        if self._ptr:
            lib().afsim_sweep_release(self._sim._h, self._ptr)
            self._ptr = C.c_void_p(None)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Simulator:
    """One afsim handle = one CUDA device + stream.  Raises when no sm_100 GPU is usable."""

    def __init__(self, device: int = 0, cuda_stream: int | None = None):
        L = lib()
        h = C.c_void_p()
        rc = L.afsim_create(int(device), C.c_void_p(cuda_stream) if cuda_stream else None, C.byref(h))
        if rc != abi.AFSIM_OK:
            raise AfsimError(rc, L.afsim_create_error().decode())
        self._h = h
        self.cuda_stream = cuda_stream  # None: the library's own stream

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().afsim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trim(self) -> None:
        """Give the handle's cached device buffers back to the driver."""
        self._check(lib().afsim_trim(self._h))

    def _check(self, rc: int) -> None:
        if rc == abi.AFSIM_OK:
            return
        msg = lib().afsim_last_error(self._h).decode()
        if rc == abi.AFSIM_INVALID_ARGUMENT:
            raise ValueError(msg)  # the reference raises PyValueError here
        raise AfsimError(rc, msg)

    def issue_peak(self, kind: int) -> float:
        """1e9 warp-lane instructions / s: kind 0 = FP64 DMUL+DADD chains, 1 = FP32 FFMA chains."""
        out = C.c_double(0.0)
        self._check(lib().afsim_measure_issue_peak(self._h, int(kind), C.byref(out)))
        return float(out.value)

    def selftest_math(self, n: int = 1 << 24) -> dict[str, int]:
        """Bit mismatches of the map kernels' device math against the CUDA library on n hashed arguments."""
        out = (C.c_uint64 * 6)()
        self._check(lib().afsim_selftest_math(self._h, int(n), out))
        return dict(zip(("log10", "exp10", "div20", "div40", "div3.75", "div_prepared"), (int(v) for v in out)))

    # ---- product resampler simulator ----
    def product_resampler(self, signals, spec: abi.AfResamplerSpec):
        """afsim_product_resampler on host buffers: signals [n_streams, n_in] f64 -> (out [n_streams, frames], shape)."""
        signals = np.ascontiguousarray(signals, dtype=np.float64)
        if signals.ndim != 2:
            raise ValueError("signals must be [n_streams, n_in]")
        n_streams, n_in = signals.shape
        shape = resampler_shape(spec, n_in, check=False)  # sizes the outputs; the call below owns the error order
        out = np.zeros((n_streams, int(shape.frames) if shape else 0), dtype=np.float64)
        f64p = C.POINTER(C.c_double)
        ins = (f64p * max(n_streams, 1))(*[signals[s].ctypes.data_as(f64p) for s in range(n_streams)])
        outs = (f64p * max(n_streams, 1))(*[out[s].ctypes.data_as(f64p) for s in range(n_streams)])
        got = abi.AfResamplerShape()
        self._check(lib().afsim_product_resampler(self._h, C.byref(spec), ins, n_streams, n_in, outs, C.byref(got)))
        return out, got

    def product_resampler_device(self, spec: abi.AfResamplerSpec, d_in: int, in_stride: int, n_streams: int, n_in: int,
                                 d_out: int, out_stride: int) -> float:
        """afsim_product_resampler_device on device pointers (torch tensors' data_ptr()); -> kernel ms."""
        ms = C.c_float(0.0)
        self._check(lib().afsim_product_resampler_device(self._h, C.byref(spec), C.c_void_p(d_in), in_stride, n_streams, n_in,
                                                         C.c_void_p(d_out), out_stride, C.byref(ms)))
        return float(ms.value)

    # ---- single-stream entry points ----
    def chain_render(self, audio, sample_rate, bands, settings: abi.AfChainSettings, return_audio: bool = False):
        audio = np.ascontiguousarray(audio, dtype=np.float32)
        m = abi.AfChainMetrics()
        out = np.zeros_like(audio) if return_audio else None
        self._check(lib().afsim_chain_render(self._h, _f32p(audio), audio.size, float(sample_rate), bands,
                                             C.byref(settings), C.byref(m), _f32p(out) if out is not None else None))
        return m, out

    def eq_render(self, audio, sample_rate, bands, return_audio: bool = False):
        audio = np.ascontiguousarray(audio, dtype=np.float32)
        st = abi.AfEqRenderStats()
        out = np.zeros_like(audio) if return_audio else None
        self._check(lib().afsim_eq_render(self._h, _f32p(audio), audio.size, float(sample_rate), bands, C.byref(st),
                                          _f32p(out) if out is not None else None))
        return st, out

    def eq_response(self, frequencies_hz, band_sets, sample_rate, typed: bool) -> np.ndarray:
        """band_sets: AfBand array holding n_sets * 10 bands -> [n_sets, n_freqs] dB."""
        freqs = np.ascontiguousarray(frequencies_hz, dtype=np.float64)
        n_sets = len(band_sets) // abi.NUM_BANDS
        out = np.zeros((n_sets, freqs.size), dtype=np.float64)
        self._check(lib().afsim_eq_response(self._h, freqs.ctypes.data_as(C.POINTER(C.c_double)), freqs.size, band_sets,
                                            n_sets, 1 if typed else 0, float(sample_rate),
                                            out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def auto_makeup_control(self, audio, sample_rate, vad_probabilities, noise_floor_db, noise_reliability,
                            settings: abi.AfAutoMakeupSettings, return_audio: bool = False):
        """simulate_auto_makeup_control -> (traces [6, ceil(n / 480)] float32 in abi.MAKEUP_TRACES order, audio | None)."""
        audio = np.ascontiguousarray(audio, dtype=np.float32)
        vad = np.ascontiguousarray(vad_probabilities if vad_probabilities is not None else [], dtype=np.float64)
        blocks = (audio.size + abi.MAKEUP_CONTROL_BLOCK - 1) // abi.MAKEUP_CONTROL_BLOCK
        traces = np.zeros((len(abi.MAKEUP_TRACES), blocks), dtype=np.float32)
        out = np.zeros_like(audio) if return_audio else None
        self._check(lib().afsim_auto_makeup_control(
            self._h, _f32p(audio), audio.size, float(sample_rate), vad.ctypes.data_as(C.POINTER(C.c_double)) if vad.size else None,
            vad.size, float(noise_floor_db), float(noise_reliability), C.byref(settings), _f32p(traces),
            _f32p(out) if out is not None else None))
        return traces, out

    def auto_makeup_sweep(self, captures, sample_rate, vads, noise_floor_db, noise_reliability, settings,
                          return_audio: bool = False):
        """Batched simulate_auto_makeup_control: one GPU pass over `captures`; vads[i] is None or an array."""
        captures = [np.ascontiguousarray(c, dtype=np.float32) for c in captures]
        n = len(captures)
        f32p, f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
        vad_arrs = [None if v is None else np.ascontiguousarray(v, dtype=np.float64) for v in vads]
        for c, v in zip(captures, vad_arrs):
            blocks = (c.size + abi.MAKEUP_CONTROL_BLOCK - 1) // abi.MAKEUP_CONTROL_BLOCK
            if v is not None and v.size != blocks:
                raise ValueError(f"expected {blocks} VAD probabilities at the 10 ms control cadence, got {v.size}")
        traces = [np.zeros((len(abi.MAKEUP_TRACES), (c.size + abi.MAKEUP_CONTROL_BLOCK - 1) // abi.MAKEUP_CONTROL_BLOCK),
                           dtype=np.float32) for c in captures]
        outs = [np.zeros_like(c) for c in captures] if return_audio else None
        m = max(n, 1)
        ptrs = (f32p * m)(*[_f32p(c) for c in captures])
        lens = (C.c_size_t * m)(*[c.size for c in captures])
        vptrs = (f64p * m)(*[v.ctypes.data_as(f64p) if v is not None and v.size else None for v in vad_arrs])
        floors = np.ascontiguousarray(noise_floor_db, dtype=np.float64)
        rels = np.ascontiguousarray(noise_reliability, dtype=np.float64)
        sets = (abi.AfAutoMakeupSettings * m)(*settings)
        tptrs = (f32p * m)(*[_f32p(t) for t in traces])
        optrs = (f32p * m)(*[_f32p(o) for o in outs]) if outs is not None else None
        self._check(lib().afsim_auto_makeup_sweep(self._h, ptrs, lens, n, float(sample_rate), vptrs,
                                                  floors.ctypes.data_as(f64p), rels.ctypes.data_as(f64p), sets, tptrs, optrs))
        return traces, outs

    # ---- sweeps ----
    @staticmethod
    def _pairs(n_passages, n_candidates, pair_passage, pair_candidate):
        if pair_passage is None:
            return None, None, n_passages * n_candidates
        pp = np.ascontiguousarray(pair_passage, dtype=np.uint32)
        pc = np.ascontiguousarray(pair_candidate, dtype=np.uint32)
        if pp.size != pc.size:
            raise ValueError("pair_passage and pair_candidate must have the same length")
        return pp, pc, pp.size

    def chain_sweep(self, passages, sample_rate, candidates, pair_passage=None, pair_candidate=None,
                    return_audio: bool = False):
        """-> (AfChainMetrics array, list of audio arrays | None); one blocking call."""
        sweep = self.prepare_sweep(passages, sample_rate, candidates, pair_passage, pair_candidate,
                                   want_audio=return_audio)
        try:
            sweep.launch()
            metrics = sweep.collect()
            audio = [sweep.collect_audio(i) for i in range(sweep.n_pairs)] if return_audio else None
        finally:
            sweep.release()
        return metrics, audio

    def prepare_sweep(self, passages, sample_rate, candidates, pair_passage=None, pair_candidate=None,
                      want_audio: bool = False) -> Sweep:
        passages = [np.ascontiguousarray(p, dtype=np.float32) for p in passages]
        f32p = C.POINTER(C.c_float)
        ptrs = (f32p * max(len(passages), 1))(*[_f32p(p) for p in passages])
        lens = (C.c_size_t * max(len(passages), 1))(*[p.size for p in passages])
        pp, pc, n_pairs = self._pairs(len(passages), len(candidates), pair_passage, pair_candidate)
        ptr = C.c_void_p()
        self._check(lib().afsim_sweep_prepare(self._h, ptrs, lens, len(passages), float(sample_rate), candidates,
                                              len(candidates), _u32p(pp), _u32p(pc), n_pairs, 1 if want_audio else 0,
                                              C.byref(ptr)))
        n_pass = max(len(passages), 1)
        pair_len = [passages[int(pp[i]) if pp is not None else i % n_pass].size for i in range(n_pairs)]
        return Sweep(self, ptr.value, n_pairs, pair_len)

    def prepare_synthetic_sweep(self, kind: int, n_passages: int, passage_len: int, sample_rate, candidates,
                                pair_passage=None, pair_candidate=None, want_audio: bool = False) -> Sweep:
        pp, pc, n_pairs = self._pairs(n_passages, len(candidates), pair_passage, pair_candidate)
        ptr = C.c_void_p()
        self._check(lib().afsim_sweep_prepare_synthetic(self._h, int(kind), n_passages, passage_len, float(sample_rate),
                                                        candidates, len(candidates), _u32p(pp), _u32p(pc), n_pairs,
                                                        1 if want_audio else 0, C.byref(ptr)))
        return Sweep(self, ptr.value, n_pairs, [passage_len] * n_pairs)


class MultiSimulator:
    """One handle for several GPUs of the box (afsim_multi_* in include/afsim.h): what the Rust host would bind.  The
    partition, the per-GPU renders and the NCCL gather of the metric structs all happen inside the library."""

    def __init__(self, device_mask: int):
        L = lib()
        h = C.c_void_p()
        rc = L.afsim_multi_create(int(device_mask), C.byref(h))
        if rc != abi.AFSIM_OK:
            raise AfsimError(rc, L.afsim_multi_create_error().decode())
        self._h = h
        self.n_devices = int(L.afsim_multi_device_count(h))
        self.last_device_ms = 0.0

    def chain_sweep(self, passages, sample_rate, candidates, pair_passage=None, pair_candidate=None):
        passages = [np.ascontiguousarray(p, dtype=np.float32) for p in passages]
        f32p = C.POINTER(C.c_float)
        ptrs = (f32p * max(len(passages), 1))(*[_f32p(p) for p in passages])
        lens = (C.c_size_t * max(len(passages), 1))(*[p.size for p in passages])
        pp, pc, n_pairs = Simulator._pairs(len(passages), len(candidates), pair_passage, pair_candidate)
        out = (abi.AfChainMetrics * max(n_pairs, 1))()
        ms = C.c_float(0.0)
        rc = lib().afsim_multi_chain_sweep(self._h, ptrs, lens, len(passages), float(sample_rate), candidates, len(candidates),
                                           _u32p(pp), _u32p(pc), n_pairs, out, C.byref(ms))
        if rc != abi.AFSIM_OK:
            msg = lib().afsim_multi_last_error(self._h).decode()
            if rc == abi.AFSIM_INVALID_ARGUMENT:
                raise ValueError(msg)
            raise AfsimError(rc, msg)
        self.last_device_ms = float(ms.value)
        return out

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().afsim_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def partition(candidates, passage_lens, pair_passage, pair_candidate, n_parts: int) -> np.ndarray:
    """afsim_multi_partition: owner of every pair (the C-side partitioner; no GPU needed)."""
    lens = (C.c_size_t * max(len(passage_lens), 1))(*[int(v) for v in passage_lens])
    pp = np.ascontiguousarray(pair_passage, dtype=np.uint32)
    pc = np.ascontiguousarray(pair_candidate, dtype=np.uint32)
    owner = np.zeros(pp.size, dtype=np.uint32)
    rc = lib().afsim_multi_partition(candidates, len(candidates), lens, len(passage_lens), _u32p(pp), _u32p(pc), pp.size,
                                     int(n_parts), _u32p(owner))
    if rc != abi.AFSIM_OK:
        raise ValueError("afsim_multi_partition: invalid arguments")
    return owner


def exported_symbols_present() -> list[str]:
    """Names from include/afsim.h that libafsim.so does not export (empty when complete)."""
    L = lib()
    return [name for name in EXPORTS if not hasattr(L, name)]


def library_loaded_path() -> str:
    lib()
    return os.fspath(LIB_PATH)
