// afsim_params.h -- data the host planner hands to the sm_100a kernels.
//
// Execution model.  A sweep is cut into BATCHES of streams (candidate x passage
// pairs) that share stage structure, length and limiter lookahead.  Inside a batch
//   * one thread renders one stream; 32 consecutive streams form a warp, so every access to the
//     time-major work buffers `buf[row][stream]` is one coalesced 128-byte transaction;
//   * the chain is cut into STAGE KERNELS (input, de-esser, EQ slices, compressor, limiter,
//     true-peak limiter + detector).  Each stage kernel advances all streams of the batch over one
//     CHUNK of samples, keeping its recurrence state in registers for the whole chunk and parking
//     it in a stream-minor state table between chunks;
//   * the stage kernels of consecutive chunks form a wavefront (stage k of chunk c only depends on
//     stage k-1 of chunk c and stage k of chunk c-1), which the host issues on one CUDA stream per
//     stage so that few-stream sweeps still fill the GPU.
//
//   CandidateParams   per candidate: every constant the recurrences need, derived on the host with
//                     the host libm exactly as the reference's constructor + setters do
//                     (rust-core/src/audio/processor/python_api.rs:400-487)
//   BatchArgs         per batch: tables and buffers (device pointers)
//   rows              [4][n_rows][S_pad] f32: the per-analysis-block rows (python_api.rs:549-557)
//   StreamAccum       per stream: running maxima / f64 square sums -> finalize kernel
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace afsim {

constexpr int kMaxSections = 40;       // 10 bands x <=4 Butterworth sections (dsp/eq.rs:32)
constexpr int kTpDelay = 20;           // dsp/true_peak.rs:11
constexpr int kMaxLookahead = 1024;    // dsp/limiter.rs:7
constexpr int kInputBlock = 480;       // block contract of the input cleanup stage (processor/tests.rs:500-549)

// per-candidate flags
enum LaneFlag : uint32_t {
    LF_EQ_FADE = 1u << 0,       // legacy band path: 10 sections fade in over F samples
    LF_DE_AUTO = 1u << 1,
    LF_C_ADAPTIVE = 1u << 2,
    LF_C_SIDECHAIN = 1u << 3,
    LF_C_AUTO_MAKEUP = 1u << 4, // compressor auto makeup (dsp/compressor.rs:598-653) over the momentary loudness meter
    LF_C_EVIDENCE = 1u << 5,    // simulate_auto_makeup_control: per-block VAD / noise evidence (dsp/compressor.rs:528-581)
};

// De-esser constants (CandidateParams::de).
enum DeField : int {
    DE_DET = 0,                 // 6 detector biquads x 5 (band b: hp at 10*b, lp at 10*b+5)
    DE_DYN_COS = 30,            // 3: cos(omega) of the dynamic EQ
    DE_DYN_ALPHA = 33,          // 3: sin(omega)/(2q)
    DE_ATTACK = 36,
    DE_RELEASE,
    DE_DET_ATTACK,
    DE_DET_RELEASE,
    DE_MAX_RED,
    DE_THRESHOLD,
    DE_RATIO_FACTOR,            // 1 - 1/ratio
    DE_RATIO_THR,               // clamp((threshold+60)*0.1, 0, 6)
    DE_TRIGGER,                 // lerp(8.0, 0.8, amount)
    DE_SLOPE,
    DE_CAP,                     // min(auto_cap, max_reduction*0.75)
    DE_CONF_FLOOR,              // clamp(lerp(0.28,0.06,amount), 0, 0.95)
    DE_BASE_FALL,
    DE_BASE_RISE,
    DE_BASE_INACTIVE,
    DE_MANUAL_CAP,              // max_reduction*0.75
    DE_FIELDS
};

struct CandidateParams {
    uint32_t flags;             // LaneFlag
    uint32_t n_sections;        // EQ sections (flattened band-major)
    float tp_ceil;              // 10f32.powf(ceiling_db as f32 / 20) clamped [1e-6, 1] (block_processor.rs:150-151)
    float tp_release;           // exp(-1/(clamp(ms,5,500)/1000*fs)) as f32 (dsp/true_peak.rs:308-313)
    float effective_ceiling_db; // python_api.rs:472-473
    uint32_t reserved;
    double eq[kMaxSections][5]; // b0 b1 b2 a1 a2 (target coefficients)
    double de[DE_FIELDS];       // DeField
    double de_det0[6][5];       // constructor coefficients of the detector biquads (fade source)
    double de_dyn0[3][5];       // dynamic EQ: constructor coefficients
    double de_dyn1[3][5];       // dynamic EQ: configured coefficients at gain 0 (fade target)
    // compressor (dsp/compressor.rs)
    double c_threshold, c_factor, c_knee, c_attack, c_det_release, c_release, c_rms, c_makeup_lin, c_sc, c_band,
        c_fast, c_charge, c_slow;
    // auto makeup (dsp/compressor.rs:598-653, 528-581): manual makeup in dB, clamped target, the three per-sample
    // smoothing coefficients (raised to the block length at every block end), evidence of the control simulator
    double c_makeup_db, c_target_lufs, c_mk_smooth, c_mk_relax, c_mk_activity;
    double c_ev_vad_reliability, c_ev_noise_floor_db, c_ev_live_reliability, c_ev_cfg_reliability;
    // limiter (dsp/limiter.rs)
    double l_ceil, l_release;
    // fixed input high-pass (audio/processor/routing.rs:826-843)
    double in_hp[5];
};

// Constants the MAP kernels need, stream-minor ([MT_FIELDS][S_pad] doubles; the float / flag fields are exactly
// representable): a map thread handles a few samples, so gathering its constants from the candidate's 2.5 KB
// CandidateParams (one 32-byte sector per lane and field) costs as much L1 / L2 traffic as its samples do; from
// this table the loads coalesce.  The serial kernels read CandidateParams once per chunk and keep doing so.
enum MapField { MT_THRESHOLD, MT_FACTOR, MT_KNEE, MT_MAKEUP_LIN, MT_L_CEIL, MT_TP_CEIL, MT_FLAGS, MT_FIELDS };
inline void fill_map_tab(double* tab, size_t stride, size_t s, const CandidateParams& p) {
    tab[MT_THRESHOLD * stride + s] = p.c_threshold;
    tab[MT_FACTOR * stride + s] = p.c_factor;
    tab[MT_KNEE * stride + s] = p.c_knee;
    tab[MT_MAKEUP_LIN * stride + s] = p.c_makeup_lin;
    tab[MT_L_CEIL * stride + s] = p.l_ceil;
    tab[MT_TP_CEIL * stride + s] = static_cast<double>(p.tp_ceil);
    tab[MT_FLAGS * stride + s] = static_cast<double>(p.flags);
}

// Stage set of a batch; every stream of a batch shares it.
enum StructureFlag : uint32_t {
    ST_DEESSER = 1u << 0,
    ST_EQ_BEFORE_DEESSER = 1u << 1,
    ST_COMPRESSOR = 1u << 2,
    ST_LIMITER = 1u << 3,
    ST_EQ = 1u << 4,
    ST_INPUT_TRUE_PEAK = 1u << 5,   // simulate_eq_v2: true-peak detector over the input as well
    ST_AUTO_MAKEUP = 1u << 6,       // compressor with auto makeup: M6 hands the gain to the serial makeup stage R7
};

// Momentary loudness meter of the auto makeup (dsp/loudness.rs over the `ebur128` crate, mode M): the folded
// 4th-order K-weighting section for the sample rate and the geometry of the 400 ms window.  The window is kept as
// `n_slots` partial sums of `slot` samples each (slot = gcd(block, window): a block always starts on a slot
// boundary), so a block end costs n_slots loads instead of a pass over the whole 400 ms ring.
struct MakeupConst {
    double b[5], a[5];
    int window;     // samples in 400 ms
    int slot;       // samples per partial sum
    int n_slots;    // window / slot
    int tail_from;  // (T mod block) mod slot: the render's final partial block leaves a slot's samples >= tail_from
};
constexpr int kMaxMakeupSub = 16;  // partial sums one block may span (block / slot)


struct StreamAccum {
    double sum_in;
    double sum_out;
    float peak_in;
    float peak_out;
    float peak_pre_tp;      // true-peak limiter input oversampler maximum
    float peak_out_tp;      // detector
    float peak_in_tp;       // ST_INPUT_TRUE_PEAK
    float limiter_gr_db;
    float tp_gr_db;
    uint32_t events;
    uint32_t non_finite;
    uint32_t reserved;
};

// Slots (doubles per stream) of the per-stage state tables.
constexpr int kStateInput = 256;  // 7 (sanitize / DC / statistics) + the adaptive cleanup stage (afsim_cleanup.h)
constexpr int kStateDeEsser = 96;
constexpr int kStateEqPerSection = 2;
constexpr int kStateCompressor = 12;
constexpr int kStateLimiter = 4;
constexpr int kStateTruePeak = 80;
constexpr int kStateMakeup = 12;

struct BatchArgs {
    const CandidateParams* params;
    const uint32_t* cand;         // [S_pad] candidate index per stream
    const uint32_t* pair;         // [S_pad] caller's pair index per stream (where its metrics go)
    const uint64_t* src_off;      // [S_pad] element offset of the stream's source signal in `signals`
    const uint64_t* audio_off;    // [S_pad] element offset of the stream's output audio in `audio`
    const float* signals;
    float* audio;                 // output audio pool, or nullptr
    float* buf_a;                 // [ring_rows][S_pad] signal up to the sample limiter's input
    float* buf_b;                 // [ring_rows][S_pad] signal after the sample limiter
    float* lim_sfx;               // [lookahead + 1][S_pad] suffix maxima of the previous limiter block (fused limiter)
    // split (R/M) path, afsim_split.h: hand-off rings shared by the serial and the map kernels
    double* w[13];                // [ring_rows][S_pad] f64 each (compressor: 4, limiter: 1, de-esser: 13)
    float* buf_c;                 // [ring_rows][S_pad] true-peak limiter output
    float* buf_p;                 // [ring_rows][S_pad] input true peaks
    double* st_input;             // [kStateInput][S_pad]
    double* st_deesser;           // [kStateDeEsser][S_pad]
    double* st_eq;                // [kStateEqPerSection * kMaxSections][S_pad]
    double* st_comp;              // [kStateCompressor][S_pad]
    double* st_lim;               // [kStateLimiter][S_pad]
    double* st_tp;                // [kStateTruePeak][S_pad]
    float* tail_hist;             // [64][S_pad] fused tail (afsim_tail.cu): last 32 rows of the limiter / true-peak limiter outputs
    float* rows;                  // [4][n_rows][S_pad]
    double* st_mk;                // [kStateMakeup][S_pad] auto makeup + loudness meter state
    double* mk_ring;              // [2][n_slots][S_pad] the meter's window: full partial sums, then tail partial sums
    float* mk_rows;               // [3][n_rows][S_pad] makeup dB / activity / reliability at block ends, or nullptr
    const double* mk_vad;         // pool of per-block VAD probabilities (LF_C_EVIDENCE), or nullptr
    const int64_t* mk_vad_off;    // [S_pad] offset of the stream's probabilities in mk_vad, -1 = no evidence
    const MakeupConst* mk_const;  // loudness meter constants of the sample rate
    StreamAccum* accum;           // [S_pad]
    // shared input stage: streams of a batch have one input stage setting, so its output depends on the passage
    // alone; it is rendered once per distinct passage (a small batch of its own) and fanned out to the streams
    const uint32_t* in_unique;    // [S_pad] index of the stream's passage among the batch's distinct ones, or nullptr
    const float* in_src;          // [ring_rows][in_stride] input stage output of the distinct passages
    const float* in_rows;         // [n_rows][in_stride] their input-level rows (row.0)
    const StreamAccum* in_accum;  // [in_stride] their input statistics
    int in_stride;
    // shared compressor front (a compressor grid over one EQ setting): sidechain signal, detector weight (dB) and
    // instantaneous peak (dB) of the distinct (passage, EQ) pairs -- what R1 + M2 produce -- read by every stream's
    // compressor instead of being recomputed; the streams then also read the EQ output itself from in_src
    // shared de-esser front (de-esser first, one detector configuration per passage): voice dB, 3 band levels dB and
    // 3 confidence targets of the distinct (passage, detector) pairs -- what R_a + M_b produce
    const double* in_de[7];       // [ring_rows][in_stride] each, or nullptr
    const double* in_det;         // [ring_rows][in_stride], or nullptr
    const double* in_wdb;
    const double* in_ipk;
    const double* eq_default;     // [10][5] constructor coefficients of the default bands (dsp/eq.rs:125-140)
    const double* de_tab;         // [DE_FIELDS][S_pad] de-esser constants, stream-minor (coalesced reads)
    const double* map_tab;        // [MT_FIELDS][S_pad] constants of the map kernels, stream-minor (MapField)
    const struct CleanupConst* cleanup;  // sample-rate constants of the adaptive input cleanup (afsim_cleanup.h)
    void* metrics;                // AfChainMetrics[n_pairs of the sweep], indexed by pair[s]
    float* fin_scratch;           // per-block finalize workspace when it does not fit shared memory
    int n_streams;                // S
    int stride;                   // S_pad
    int n_samples;                // T (uniform in the batch)
    int ring_rows;                // rows of buf_a / buf_b (a multiple of the chunk, >= chunk + lookahead + 1)
    int n_rows;                   // analysis blocks per stream
    int n_pad;                    // power of two >= n_rows (finalize sort)
    int block_samples;            // analysis block (round(fs*0.020), python_api.rs:512-513)
    int fade_samples;             // biquad crossfade length F (dsp/biquad.rs:12-19)
    int lookahead;                // limiter lookahead L in samples
    int input_stage;              // AfInputStage
    int stage_inputs;             // 1: serial kernels stage their inputs through shared memory (few-stream batches)
    uint32_t structure;           // StructureFlag
    int map_blocks_per_sm;        // AFSIM_MAP_BLOCKS_PER_SM: grid cap of the map kernels (0 = none)
    int fir_blocks_per_sm;        // AFSIM_FIR_BLOCKS_PER_SM: grid cap of the FIR / limiter-window maps (-1 = the default cap)
};

}  // namespace afsim
