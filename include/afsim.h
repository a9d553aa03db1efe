/*
 * afsim.h -- C ABI of the B200-native batched chain simulator (libafsim.so).
 *
 * This is the drop-in boundary for ONE path of FueledByRedBull/audio-forge: the
 * native offline chain simulator.  Every entry point below names the reference
 * interface it replaces (paths relative to the reference checkout).  The Rust
 * host would bind these with `extern "C"` / `libloading`, exactly as it already
 * binds its DeepFilterNet plugin (rust-core/src/dsp/deepfilter_ffi.rs:164-176,
 * 335-386: opaque handle create/free, caller-owned buffers, integer status).
 * INTEGRATION.md shows that binding and the ctypes shim used in this repo.
 *
 * Conventions
 *   - plain C, POD structs only (all `#[repr(C)]`-compatible), no torch types;
 *   - the caller owns every buffer; the library owns only the handle (CUDA
 *     context objects, stream, scratch);
 *   - every call is synchronous (work is queued on the handle's stream and the
 *     stream is synchronised before returning) unless it says "async";
 *   - return value: AfStatus; afsim_last_error(handle) gives the message;
 *   - there is NO CPU fallback: without a CUDA device afsim_create fails.
 */
#ifndef AFSIM_H
#define AFSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFSIM_NUM_BANDS 10
#define AFSIM_ABI_VERSION 1

typedef enum AfStatus {
    AFSIM_OK = 0,
    AFSIM_INVALID_ARGUMENT = 1, /* reference: PyValueError (python_api.rs:388-398, lib.rs:105-141,160-186) */
    AFSIM_CUDA_ERROR = 2,
    AFSIM_OUT_OF_MEMORY = 3,
    AFSIM_UNSUPPORTED = 4 /* a settings combination this build does not run on the GPU (fails loudly) */
} AfStatus;

/* Stable filter ids, reference rust-core/src/dsp/eq.rs:46-53 (EqFilterType). */
typedef enum AfFilterType {
    AF_LOW_SHELF = 0,
    AF_BELL = 1,
    AF_HIGH_SHELF = 2,
    AF_NOTCH = 3,
    AF_HIGH_PASS = 4,
    AF_LOW_PASS = 5
} AfFilterType;

/* Input stage in front of the chain.  The reference's offline simulator has none
 * (AF_INPUT_NONE reproduces it exactly); the other values add the live loop's
 * input stage (rust-core/src/audio/processor/routing.rs:826-843 DC block + 80 Hz
 * high-pass, and routing.rs:213-609 adaptive hum/rumble cleanup, block = 480
 * samples as in processor/tests.rs:500-549). */
typedef enum AfInputStage {
    AF_INPUT_NONE = 0,
    AF_INPUT_DC_HP80 = 1,       /* cleanup "off": DC block then fixed 80 Hz HP */
    AF_INPUT_CLEANUP_GENTLE = 2,
    AF_INPUT_CLEANUP_STRONG = 3
} AfInputStage;

/* One EQ band.  Legacy 3-tuple callers (python_api.rs:383, lib.rs:100-104) fill
 * frequency/gain/q only; typed callers (lib.rs:152 PyEqBandV2) fill everything. */
typedef struct AfBand {
    double frequency_hz;
    double gain_db;
    double q;
    uint8_t filter_type;          /* AfFilterType */
    uint8_t slope_db_per_octave;  /* 12, 24, 36 or 48 (pass filters) */
    uint8_t enabled;
    uint8_t reserved[5];
} AfBand;

/* The flat settings dict of simulate_auto_eq_chain (python_api.rs:415-487; the
 * keys headroom.py:48-84 sends, plus eq_before_deesser, limiter_lookahead_ms and
 * eq_bands_v2 presence = use_typed_bands).  afsim_chain_settings_default() fills
 * the reference defaults. */
typedef struct AfChainSettings {
    uint8_t use_typed_bands;      /* 1: eq_bands_v2 branch (set_band_config + reset) ; 0: legacy setters (72-sample fade-in) */
    uint8_t eq_before_deesser;
    uint8_t deesser_enabled;
    uint8_t deesser_auto_enabled;
    uint8_t compressor_enabled;
    uint8_t compressor_adaptive_release;
    uint8_t compressor_auto_makeup_enabled;
    uint8_t compressor_sidechain_highpass_enabled;
    uint8_t limiter_enabled;
    uint8_t limiter_careful_output_enabled;
    uint8_t input_stage;          /* AfInputStage; 0 = reference offline behaviour */
    uint8_t reserved[5];
    double deesser_auto_amount;
    double deesser_low_cut_hz;
    double deesser_high_cut_hz;
    double deesser_threshold_db;
    double deesser_ratio;
    double deesser_attack_ms;
    double deesser_release_ms;
    double deesser_max_reduction_db;
    double compressor_threshold_db;
    double compressor_ratio;
    double compressor_attack_ms;
    double compressor_release_ms;
    double compressor_makeup_gain_db;
    double compressor_base_release_ms;
    double compressor_target_lufs;
    double limiter_ceiling_db;
    double limiter_release_ms;
    double limiter_lookahead_ms;
} AfChainSettings;

/* One candidate of a sweep = 10 bands + the chain settings. */
typedef struct AfCandidate {
    AfBand bands[AFSIM_NUM_BANDS];
    AfChainSettings settings;
} AfCandidate;

/* The dict simulate_auto_eq_chain returns (python_api.rs:649-713), 1:1. */
typedef struct AfChainMetrics {
    float input_sample_peak_db;
    float input_rms_db;
    float output_sample_peak_db;
    float pre_limiter_true_peak_db;
    float output_true_peak_db;
    float output_rms_db;
    float limiter_effective_ceiling_db;
    float sample_headroom_db;
    float pre_limiter_true_peak_headroom_db;
    float true_peak_headroom_db;
    float limiter_gain_reduction_db;
    float true_peak_limiter_gain_reduction_db;
    float compressor_gain_reduction_db;
    float deesser_gain_reduction_db;
    float compressor_gain_reduction_median_db;
    float compressor_gain_reduction_p95_db;
    float compressor_gain_reduction_active_ratio;
    float active_output_gain_db;
    float silence_output_gain_db;
    float silence_level_delta_db;
    float compressor_pumping_score_db;
    float deesser_gain_reduction_median_db;
    float deesser_gain_reduction_p95_db;
    float analysis_block_ms;
    float active_analysis_threshold_db;
    uint32_t non_finite_output;
    uint64_t true_peak_limited_events;
    uint64_t active_analysis_block_count;
    uint64_t processed_samples;
    double candidate_runtime_ms; /* wall time of the call divided over its streams */
} AfChainMetrics;

/* The dict simulate_eq_v2 returns (lib.rs:269-286). */
typedef struct AfEqRenderStats {
    float input_sample_peak;
    float output_sample_peak;
    float input_true_peak;
    float output_true_peak;
    double input_rms;
    double output_rms;
    double max_response_db;
    double runtime_ms;
    uint64_t sample_count;
    uint64_t algorithmic_latency_samples;
    uint32_t non_finite_output;
    uint32_t reserved;
} AfEqRenderStats;

/* settings dict of simulate_auto_makeup_control (python_api.rs:168-192); afsim_auto_makeup_settings_default()
 * fills the reference defaults (-24 dB, 3:1, 10 ms, 180 ms, 0 dB, -18 LUFS, adaptive release on, sidechain
 * high-pass on, vad_reliability 1). */
typedef struct AfAutoMakeupSettings {
    double threshold_db;
    double ratio;
    double attack_ms;
    double release_ms;
    double makeup_gain_db;
    double target_lufs;
    double vad_reliability;
    uint8_t adaptive_release;
    uint8_t sidechain_highpass_enabled;
    uint8_t reserved[6];
} AfAutoMakeupSettings;

#define AFSIM_MAKEUP_CONTROL_BLOCK 480 /* CONTROL_BLOCK_SIZE, python_api.rs:135 */
#define AFSIM_MAKEUP_TRACES 6          /* makeup_gain_db, activity, reliability, gain_reduction_db, input_rms_db, output_rms_db */

typedef struct AfsimHandle AfsimHandle;   /* opaque */
typedef struct AfsimSweep AfsimSweep;     /* opaque: a sweep resident in HBM */

/* ---- lifetime ------------------------------------------------------------ */

int afsim_abi_version(void);

/* Create a simulator bound to CUDA device `device_ordinal`.  `cuda_stream` may
 * be NULL (the library creates its own stream) or an existing cudaStream_t to
 * launch on (so a caller's CUDA events see the work).  Fails (AFSIM_CUDA_ERROR)
 * when no usable sm_100 device is present: there is no CPU path. */
int afsim_create(int device_ordinal, void* cuda_stream, AfsimHandle** out_handle);
/* Threading: calls on ONE handle are serialised by the library (a per-handle lock; the reference's pyfunctions hold
 * the GIL for the whole render, python_api.rs:378); different handles are independent.
 * Lifetime: afsim_destroy may be called while sweeps prepared on the handle are still alive -- it waits for the
 * device, detaches them (they keep their device memory and stay valid for afsim_sweep_release only), and every
 * other afsim_sweep_* call on a detached sweep is an error the caller must not make.  afsim_sweep_release ignores
 * its handle argument for a detached sweep, so release-after-destroy is safe in either order. */
void afsim_destroy(AfsimHandle* handle);
/* Gives the handle's cache of freed device buffers (sweep rings are recycled between calls; at most
 * AFSIM_POOL_MAX_GB, default 96) back to the driver. */
int afsim_trim(AfsimHandle* handle);
/* Message of the last failure on this handle ("" if none).  Valid until the next call. */
const char* afsim_last_error(const AfsimHandle* handle);
/* Message of the last afsim_create failure in this thread. */
const char* afsim_create_error(void);

/* Reference defaults of every settings key (python_api.rs:415-487). */
void afsim_chain_settings_default(AfChainSettings* out);
/* Reference default band layout (dsp/eq.rs:11-23,127-140). */
void afsim_default_bands(AfBand out[AFSIM_NUM_BANDS]);

/* ---- single-stream entry points (the reference's pyfunctions) ------------ */

/* Replaces simulate_auto_eq_chain (python_api.rs:378-714).  `out_audio` is NULL
 * or n floats (settings["return_output_audio"]). */
int afsim_chain_render(AfsimHandle* handle, const float* audio, size_t n, double sample_rate,
                       const AfBand bands[AFSIM_NUM_BANDS], const AfChainSettings* settings,
                       AfChainMetrics* out_metrics, float* out_audio);

/* Replaces simulate_eq_v2 (lib.rs:214-288).  Non-finite input is rejected. */
int afsim_eq_render(AfsimHandle* handle, const float* audio, size_t n, double sample_rate,
                    const AfBand bands[AFSIM_NUM_BANDS], AfEqRenderStats* out_stats,
                    float* out_audio);

/* Replaces eq_magnitude_response (typed=0, lib.rs:99-150) and
 * eq_magnitude_response_v2 (typed=1, lib.rs:191-212) for `n_sets` band sets at
 * once: out_db[set * n_freqs + i]. */
int afsim_eq_response(AfsimHandle* handle, const double* frequencies_hz, size_t n_freqs,
                      const AfBand* bands /* n_sets * 10 */, size_t n_sets, int typed,
                      double sample_rate, double* out_db);

/* Replaces simulate_auto_makeup_control (python_api.rs:118-276): streams one capture through the compressor's
 * auto-makeup controller in 480-sample control blocks.  vad_probabilities: NULL / n_vad = 0 (no evidence), or
 * exactly ceil(n / 480) values in [0, 1].  out_traces: AFSIM_MAKEUP_TRACES arrays of ceil(n / 480) floats,
 * trace-major, in the order named above.  out_audio: NULL or n floats.  The loudness meter restates the
 * third-party `ebur128` crate (momentary mode); see DESIGN.md for its parity status. */
void afsim_auto_makeup_settings_default(AfAutoMakeupSettings* out);
int afsim_auto_makeup_control(AfsimHandle* handle, const float* audio, size_t n, double sample_rate,
                              const double* vad_probabilities, size_t n_vad, double noise_floor_db,
                              double noise_reliability, const AfAutoMakeupSettings* settings, float* out_traces,
                              float* out_audio);
/* The same for `n_streams` captures at once (one GPU pass): capture i has len[i] samples, settings[i], and either
 * no evidence (vad[i] == NULL) or ceil(len[i] / 480) probabilities.  out_traces[i]: 6 x ceil(len[i] / 480)
 * floats; out_audio: NULL or n_streams pointers (each NULL or len[i] floats). */
int afsim_auto_makeup_sweep(AfsimHandle* handle, const float* const* audio, const size_t* len, size_t n_streams,
                            double sample_rate, const double* const* vad, const double* noise_floor_db,
                            const double* noise_reliability, const AfAutoMakeupSettings* settings,
                            float* const* out_traces, float* const* out_audio);

/* ---- batched sweeps ------------------------------------------------------ */

/* Candidate x passage sweep: what apply_headroom_validation (headroom.py:292-354)
 * and _calibrate_compressor_threshold (voice_setup.py:742-1079) do one call at a
 * time.  pair_passage/pair_candidate (n_pairs each) select the streams; both NULL
 * means the full cross product, candidate-major: pair = candidate * n_passages +
 * passage.  out_metrics has n_pairs entries.  out_audio is NULL or n_pairs
 * pointers (each NULL or passage_len floats). */
int afsim_chain_sweep(AfsimHandle* handle, const float* const* passages, const size_t* passage_len,
                      size_t n_passages, double sample_rate, const AfCandidate* candidates,
                      size_t n_candidates, const uint32_t* pair_passage,
                      const uint32_t* pair_candidate, size_t n_pairs, AfChainMetrics* out_metrics,
                      float* const* out_audio);

/* The same sweep split into upload / run / download so that a caller can keep
 * the inputs resident in HBM and re-run (bench.py's kernel-only number). */
int afsim_sweep_prepare(AfsimHandle* handle, const float* const* passages,
                        const size_t* passage_len, size_t n_passages, double sample_rate,
                        const AfCandidate* candidates, size_t n_candidates,
                        const uint32_t* pair_passage, const uint32_t* pair_candidate,
                        size_t n_pairs, int want_audio, AfsimSweep** out_sweep);
/* Device-side synthetic passages (bench only: avoids a 94 GB host buffer for the
 * batch true-peak config).  kind 0 = speech-like, 1 = hot white noise; seeds as
 * SURVEY 8(d). */
int afsim_sweep_prepare_synthetic(AfsimHandle* handle, int kind, size_t n_passages,
                                  size_t passage_len, double sample_rate,
                                  const AfCandidate* candidates, size_t n_candidates,
                                  const uint32_t* pair_passage, const uint32_t* pair_candidate,
                                  size_t n_pairs, int want_audio, AfsimSweep** out_sweep);
/* async: queues the render + reduction kernels on the handle's stream. */
int afsim_sweep_launch(AfsimHandle* handle, AfsimSweep* sweep);
/* Waits for the stream and copies n_pairs metrics to the host. */
int afsim_sweep_collect(AfsimHandle* handle, AfsimSweep* sweep, AfChainMetrics* out_metrics);
/* Copies one stream's rendered audio (want_audio sweeps) to the host. */
int afsim_sweep_collect_audio(AfsimHandle* handle, AfsimSweep* sweep, size_t pair, float* out_audio,
                              size_t n);
/* Waits for the stream and reports an asynchronous failure of the last launch (AFSIM_OK otherwise) without copying
 * anything: for callers that read the metrics through afsim_sweep_metrics_device_ptr instead of afsim_sweep_collect. */
int afsim_sweep_status(AfsimHandle* handle, AfsimSweep* sweep);
/* Device pointer of the n_pairs AfChainMetrics (for an NCCL gather without a host hop). */
void* afsim_sweep_metrics_device_ptr(AfsimSweep* sweep);
/* Kernels launched by one afsim_sweep_launch. */
int afsim_sweep_kernel_count(const AfsimSweep* sweep);
/* Milliseconds the render kernel(s) of the last afsim_sweep_launch took (CUDA
 * events on the handle's stream); waits for completion. */
int afsim_sweep_last_render_ms(AfsimHandle* handle, AfsimSweep* sweep, float* out_ms);
void afsim_sweep_release(AfsimHandle* handle, AfsimSweep* sweep);

/* ---- one handle for all GPUs of the box --------------------- */

/* The Rust host binds ONE handle per process (the shape of its runtime-loaded plugin, dsp/deepfilter_ffi.rs:335-386).
 * afsim_multi_create opens one simulator per GPU named in `device_mask` (bit d = CUDA device d) and, for more than one
 * GPU, one NCCL communicator over them (libnccl.so.2 is resolved at run time; AFSIM_UNSUPPORTED when it is missing).
 * afsim_multi_chain_sweep is afsim_chain_sweep over all of them: the candidate x passage streams are partitioned by
 * cost (afsim_multi_partition), every GPU renders its shard (passages are replicated), and the per-stream metric
 * structs are all-gathered device to device with ncclAllGather -- the only exchange on the path; out_metrics is in the
 * caller's pair order.  out_device_ms (nullable): device time of the slowest GPU, render + gather (CUDA events). */
typedef struct AfsimMulti AfsimMulti;
int afsim_multi_create(uint32_t device_mask, AfsimMulti** out_handle);
void afsim_multi_destroy(AfsimMulti* handle);
int afsim_multi_device_count(const AfsimMulti* handle);
const char* afsim_multi_last_error(const AfsimMulti* handle);
const char* afsim_multi_create_error(void);
int afsim_multi_chain_sweep(AfsimMulti* handle, const float* const* passages, const size_t* passage_len,
                            size_t n_passages, double sample_rate, const AfCandidate* candidates, size_t n_candidates,
                            const uint32_t* pair_passage, const uint32_t* pair_candidate, size_t n_pairs,
                            AfChainMetrics* out_metrics, float* out_device_ms);
/* The partition afsim_multi_chain_sweep uses (deterministic; also what audio_forge_b200/sharding.py computes for the
 * one-process-per-GPU launch): out_owner[i] = part of pair i, parts balanced by (40 + EQ sections) x samples. */
int afsim_multi_partition(const AfCandidate* candidates, size_t n_candidates, const size_t* passage_len,
                          size_t n_passages, const uint32_t* pair_passage, const uint32_t* pair_candidate,
                          size_t n_pairs, int n_parts, uint32_t* out_owner);

/* ---- product resampler simulator ---------------------------- */

/* Replaces simulate_product_resampler (rust-core/src/audio/processor/resampling.rs:170-262): the signal through
 * rubato 0.14's SincFixedIn<f64> -- cubic interpolation between 256 windowed-sinc phases, `chunk_size` blocks, one
 * zero-padded partial block, zero-input flush blocks until expected + delay frames exist -- for `n_streams` signals of
 * equal length at once.  Every frame is rendered where the reference's block loop puts it (the loop is walked on the
 * host, addition for addition); results agree with the crate to f64 rounding of the dot products.  rubato is not
 * vendored in the reference tree: `calculate_cutoff` is available only for the configurations the reference ships and
 * evaluates (128 blackman -- the product default, resampling.rs:131-138 --, 128 and 256 blackman_harris_squared;
 * python/tools/evaluate_resampler_quality.py), other pairs return AFSIM_UNSUPPORTED.  The reference's fourth result,
 * wall-clock nanoseconds per block, is a host measurement and has no counterpart here. */
typedef enum AfResamplerWindow { /* resampler_window_from_name, resampling.rs:158-168 */
    AF_WINDOW_BLACKMAN_HARRIS = 0,
    AF_WINDOW_BLACKMAN_HARRIS2 = 1,
    AF_WINDOW_BLACKMAN = 2,
    AF_WINDOW_BLACKMAN2 = 3,
    AF_WINDOW_HANN = 4,
    AF_WINDOW_HANN2 = 5
} AfResamplerWindow;
typedef struct AfResamplerSpec {
    uint32_t input_rate, output_rate;
    uint32_t chunk_size; /* 1 .. 1024 (RESAMPLER_CHUNK_SIZE) */
    uint32_t sinc_len;   /* power of two, 32 .. 2048 */
    int32_t window;      /* AfResamplerWindow */
} AfResamplerSpec;
typedef struct AfResamplerShape {
    uint64_t frames;          /* frames the block loop produces (>= expected_frames + delay): length of every output */
    uint64_t expected_frames; /* round(n_in * output_rate / input_rate) */
    uint32_t delay;           /* output_delay() of the crate, in output frames */
    uint32_t blocks;          /* process calls the reference makes (full + partial + flush) */
} AfResamplerShape;
/* product_resampler_configuration (resampling.rs:263-272): 128 taps, blackman, chunk 1024; rates left 0. */
void afsim_resampler_spec_default(AfResamplerSpec* out);
/* Validation (the reference's messages) and output shape; host only, needs no device.  `err` (nullable) receives the
 * message on failure. */
int afsim_product_resampler_shape(const AfResamplerSpec* spec, size_t n_in, AfResamplerShape* out_shape, char* err,
                                  size_t err_capacity);
/* Host buffers: samples[s] = n_in doubles (finite, else AFSIM_INVALID_ARGUMENT "samples must be finite"), out[s] =
 * shape.frames doubles. */
int afsim_product_resampler(AfsimHandle* handle, const AfResamplerSpec* spec, const double* const* samples,
                            size_t n_streams, size_t n_in, double* const* out, AfResamplerShape* out_shape);
/* The same on device-resident buffers (d_in[s * in_stride + i], d_out[s * out_stride + f]); asynchronous on the
 * handle's stream up to the final synchronisation; out_ms (nullable) = kernel time by CUDA events.  Non-finite input is
 * the caller's responsibility here. */
int afsim_product_resampler_device(AfsimHandle* handle, const AfResamplerSpec* spec, const double* d_in,
                                   size_t in_stride, size_t n_streams, size_t n_in, double* d_out, size_t out_stride,
                                   float* out_ms);
/* Test / audit hook: the planner's phase table ([256][sinc_len], nullable) and frame list (base sample, phase,
 * cubic abscissa; each nullable, shape.frames entries) so that the CPU tier can hold them against the oracle. */
int afsim_product_resampler_plan(const AfResamplerSpec* spec, size_t n_in, double* out_table, int64_t* out_base,
                                 int32_t* out_phase, double* out_frac);

/* ---- measurement helpers (bench.py; no reference counterpart) ------------------ */

/* Stage kinds reported by afsim_sweep_profile_stages. */
typedef enum AfStageKind {
    AF_STAGE_INPUT = 0,
    AF_STAGE_INPUT_TRUE_PEAK = 1,
    AF_STAGE_DEESSER = 2,
    AF_STAGE_EQ = 3,
    AF_STAGE_COMPRESSOR = 4,
    AF_STAGE_LIMITER = 5,
    AF_STAGE_OUTPUT = 6, /* true-peak limiter + detector + output statistics */
    AF_STAGE_FINALIZE = 7,
    /* split (serial recurrence R / parallel map M) kernels of few-stream batches, in chain order:
     * compressor R1 M2 R3 M4 R5 M6, limiter M R, true-peak FIR-in R FIR-out */
    AF_STAGE_SPLIT_BASE = 8,
    /* ... AF_STAGE_SPLIT_BASE + 14 = auto makeup R7 */
    AF_STAGE_INPUT_FANOUT = 40, /* copy of the shared input stage (one render per distinct passage) to every stream */
    AF_STAGE_TAIL = 41 /* limiter -> true-peak limiter -> detector + output statistics fused in one SM-local, TMA-fed kernel */
} AfStageKind;

/* Runs the first batch of the sweep once with the stage kernels SERIALISED on the handle's stream
 * and CUDA events around every launch of the first `max_chunks` chunks (0 = all).  Per stage of the
 * chain (in chain order; EQ slices are separate entries): out_kind[i] = AfStageKind, out_ms[i] = sum
 * of its launch durations, out_launches[i] = launches timed.  *out_n = entries written (<= capacity). */
int afsim_sweep_profile_stages(AfsimHandle* handle, AfsimSweep* sweep, int max_chunks, int capacity,
                               int* out_kind, float* out_ms, int* out_launches, int* out_n);
/* Shape of what afsim_sweep_profile_stages / _wavefront time: out_info = {batches of the sweep, streams of the first
 * batch, samples per chunk, ring slots, stages, samples per stream}; out_stage_streams[i] = streams ONE launch of stage
 * i really processes (a shared-prefix stage runs on the distinct (passage, setting) pairs only; a cut compressor grid's
 * first batch is one piece).  The roofline arithmetic of bench.py uses these, not the sweep's pair count. */
int afsim_sweep_batch_info(const AfsimSweep* sweep, int capacity, int out_info[6], int* out_stage_streams, int* out_n);
/* The same batch in the LIVE wavefront (every stage on its own stream, as afsim_sweep_launch runs it) with timing
 * events around the launches of chunks [first_chunk, first_chunk + n_chunks): per stage, out_busy_ms[i] = mean time
 * from "launch eligible" (its waits satisfied) to "kernel done" -- the stage's duration under contention -- and
 * out_period_ms[i] = mean time between the completions of consecutive chunks on that stage (the pipeline period). */
int afsim_sweep_profile_wavefront(AfsimHandle* handle, AfsimSweep* sweep, int first_chunk, int n_chunks, int capacity,
                                  int* out_kind, float* out_busy_ms, float* out_period_ms, int* out_n);
/* Self-test of the device math of the map kernels (csrc/afsim_math.h: the CUDA library's log10 / exp10 algorithms
 * with constant-bank coefficients, division by a constant as multiply + exact remainder + correction) against the
 * library routines / the division instruction sequence on n hashed arguments: out_mismatches[k] = results that
 * differ in any bit, k = 0 log10, 1 exp10, 2 x/20, 3 x/40, 4 x/3.75, 5 a/b with a shared prepared divisor.  All
 * six must be 0. */
int afsim_selftest_math(AfsimHandle* handle, uint64_t n, uint64_t out_mismatches[6]);
/* Sustained issue rate of this GPU in 1e9 warp-lane instructions per second: kind 0 = dependent
 * FP64 DMUL+DADD chains (the instruction mix of the unfused biquads), kind 1 = FP32 FFMA chains. */
int afsim_measure_issue_peak(AfsimHandle* handle, int kind, double* out_giga_instr_per_s);

#ifdef __cplusplus
}
#endif
#endif /* AFSIM_H */
