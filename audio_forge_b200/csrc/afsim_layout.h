// afsim_layout.h -- HBM layout shared by the host planner and the sm_100a kernels.
//
// One WARP renders 32 streams (one candidate x passage pair per lane).  Everything a lane
// needs lives in lane-interleaved tables, `table[(group * FIELDS + field) * 32 + lane]`, so a
// warp's access to one field is one coalesced 128/256-byte transaction:
//   params  (f64)  per-stream constants derived on the host (coefficients, thresholds ...)
//   state   (f64)  per-stream recurrence state parked between time tiles
//   fstate  (f32)  per-stream f32 state (true-peak FIR windows, limiter gains ...)
//   rows    (f32)  per-stream per-analysis-block rows [stream][4][rows_pitch]
//   accum   (f32/f64) per-stream running maxima / sums -> finalize kernel
// "group" = warp index within the launch.
#pragma once
#include <stdint.h>

namespace afsim {

constexpr int kLanes = 32;
constexpr int kMaxSections = 40;       // 10 bands x <=4 Butterworth sections (dsp/eq.rs:32)
constexpr int kTile = 32;              // time tile (samples) a warp processes stage by stage
constexpr int kFirTaps = 32;
constexpr int kTpDelay = 20;           // dsp/true_peak.rs:11
constexpr int kMaxLookahead = 1024;    // dsp/limiter.rs:7

// ---- f64 parameter fields ---------------------------------------------------------------------------
enum ParamField : int {
    // EQ: target coefficients of section s at P_EQ + 5*s + {b0,b1,b2,a1,a2}
    P_EQ = 0,
    // de-esser (dsp/deesser.rs): 6 detector biquads (band b: hp at P_DE_DET + 10*b, lp at +5),
    P_DE_DET = P_EQ + 5 * kMaxSections,
    // pre-fade (constructor) coefficients of the 6 detector biquads and 3 dynamic EQs
    P_DE_DET0 = P_DE_DET + 30,
    P_DE_DYN0 = P_DE_DET0 + 30,         // 3 x 5 active coefficients at t=0 (default bounds, gain 0)
    P_DE_DYN1 = P_DE_DYN0 + 15,         // 3 x 5 pending coefficients (final bounds, gain 0)
    P_DE_DYN_COS = P_DE_DYN1 + 15,      // 3: cos(omega) of the final dynamic EQ
    P_DE_DYN_ALPHA = P_DE_DYN_COS + 3,  // 3: sin(omega)/(2q)
    P_DE_ATTACK = P_DE_DYN_ALPHA + 3,
    P_DE_RELEASE,
    P_DE_DET_ATTACK,
    P_DE_DET_RELEASE,
    P_DE_MAX_RED,
    P_DE_THRESHOLD,
    P_DE_RATIO_FACTOR,                  // 1 - 1/ratio
    P_DE_RATIO_THR,                     // clamp((threshold+60)*0.1, 0, 6)
    P_DE_TRIGGER,                       // lerp(8.0, 0.8, amount)
    P_DE_SLOPE,
    P_DE_CAP,                           // min(auto_cap, max_reduction*0.75)
    P_DE_CONF_FLOOR,                    // clamp(lerp(0.28,0.06,amount), 0, 0.95)
    P_DE_BASE_FALL,
    P_DE_BASE_RISE,
    P_DE_BASE_INACTIVE,
    // compressor (dsp/compressor.rs)
    P_C_THRESHOLD,
    P_C_FACTOR,                         // 1 - 1/ratio
    P_C_KNEE,
    P_C_ATTACK,
    P_C_DET_RELEASE,
    P_C_RELEASE,                        // GR smoothing release (non-adaptive: base_release quirk)
    P_C_RMS,
    P_C_MAKEUP_LIN,                     // 10^(makeup/20) (constant when auto-makeup is off)
    P_C_SC_COEFF,
    P_C_BAND,
    P_C_FAST,
    P_C_CHARGE,
    P_C_SLOW,
    // limiter (dsp/limiter.rs)
    P_L_CEIL,
    P_L_RELEASE,
    // input stage: fixed 80 Hz high-pass (routing.rs:826-843)
    P_IN_HP,                            // 5 coefficients
    P_COUNT = P_IN_HP + 5
};

// ---- f32 parameter fields ---------------------------------------------------------------------------
enum FParamField : int {
    FP_TP_CEIL = 0,   // 10f32.powf(ceiling_db as f32 / 20) clamped [1e-6, 1]
    FP_TP_RELEASE,    // (exp(..) as f32)
    FP_COUNT
};

// ---- per-lane integer flags ---------------------------------------------------------------------------
enum LaneFlag : uint32_t {
    LF_ACTIVE = 1u << 0,        // lane carries a real stream
    LF_EQ_FADE = 1u << 1,       // legacy band path: 10 sections fade in over F samples
    LF_DE_AUTO = 1u << 2,
    LF_C_ADAPTIVE = 1u << 3,
    LF_C_SIDECHAIN = 1u << 4,
};

// ---- f64 state fields ---------------------------------------------------------------------------------
enum StateField : int {
    S_EQ = 0,                           // section s: z1,z2,pz1,pz2 at S_EQ + 4*s
    S_DE_DET = S_EQ + 4 * kMaxSections, // 6 detector biquads x {z1,z2,pz1,pz2}
    S_DE_DYN = S_DE_DET + 24,           // 3 dynamic EQs x {z1,z2,pz1,pz2}
    S_DE_DYN_COEF = S_DE_DYN + 12,      // 3 x 5 live coefficients of the dynamic EQs
    S_DE_DYN_GAIN = S_DE_DYN_COEF + 15, // 3: last gain_db the dynamic EQ was built for
    S_DE_DYN_CANCEL = S_DE_DYN_GAIN + 3, // 3: 1.0 once set_gain_db_immediate cancelled the fade
    S_DE_ENV = S_DE_DYN_CANCEL + 3,     // 3
    S_DE_CONF = S_DE_ENV + 3,           // 3
    S_DE_BASE = S_DE_CONF + 3,          // 3
    S_DE_RED = S_DE_BASE + 3,           // 3
    S_DE_BROADBAND = S_DE_RED + 3,
    S_DE_CURRENT,                       // current_reduction_db
    S_C_PREV_IN,
    S_C_PREV_OUT,
    S_C_LOW,
    S_C_VOICED,
    S_C_PRESENCE,
    S_C_PEAK_ENV,
    S_C_RMS_ENV,
    S_C_GR,
    S_C_FAST,
    S_C_SLOW,
    S_L_GAIN,
    S_L_MIN_GAIN,                       // min gain over the whole render (-> peak gain reduction)
    S_IN_HP,                            // z1,z2 of the fixed input high-pass
    S_SUM_IN = S_IN_HP + 2,             // f64 square sums (python_api.rs:490-491)
    S_SUM_OUT,
    S_BLK_IN,                           // per-block partial sums
    S_BLK_OUT,
    S_COUNT
};

// ---- f32 state fields ---------------------------------------------------------------------------------
enum FStateField : int {
    FS_WIN_IN = 0,                      // 31 samples of true-peak-limiter input history (oldest first)
    FS_WIN_OUT = FS_WIN_IN + 31,        // 31 samples of output history for the detector
    FS_TP_GAIN = FS_WIN_OUT + 31,
    FS_TP_MIN_GAIN,
    FS_TP_BLOCK_LIMITED,                // 1.0 when the current block had a new attack
    FS_PEAK_IN,
    FS_PEAK_OUT,
    FS_PEAK_PRE_TP,
    FS_PEAK_OUT_TP,
    FS_MAX_COMP_GR,
    FS_MAX_DE_GR,
    FS_EVENTS,
    FS_NONFINITE,
    FS_DC_X1,
    FS_DC_Y1,
    FS_TILE_MAX,                        // ring of per-32-sample maxima for the limiter window: kTileMaxSlots
    FS_COUNT = FS_TILE_MAX + 40
};
constexpr int kTileMaxSlots = 40;       // >= kMaxLookahead/32 + 3

// ---- per-group (warp) header ----------------------------------------------------------------------------
struct GroupHeader {
    uint32_t n_samples;        // stream length T (uniform in the group)
    uint32_t block_samples;    // analysis block (round(fs*0.020), python_api.rs:512-513)
    uint32_t fade_samples;     // biquad crossfade length F (dsp/biquad.rs:12-19)
    uint32_t lookahead;        // limiter lookahead L in samples (uniform in the group)
    uint32_t max_sections;     // max EQ sections over the lanes
    uint32_t flags;            // GroupFlag
    uint32_t first_stream;     // slot of lane 0 (= 32 * group; padding lanes own scratch slots)
    uint32_t n_streams;        // real streams in this group (<= 32)
    uint32_t rows_pitch;       // row capacity per stream
    uint32_t reserved[3];
    uint64_t src_offset[kLanes];   // element offset of each lane's source signal in the signal pool
    uint64_t audio_offset[kLanes]; // element offset of each lane's output audio (GF_WRITE_AUDIO)
    uint32_t lane_flags[kLanes];
    uint32_t n_sections[kLanes];
};

enum GroupFlag : uint32_t {
    GF_DEESSER = 1u << 0,
    GF_EQ = 1u << 1,
    GF_EQ_BEFORE_DEESSER = 1u << 2,
    GF_COMPRESSOR = 1u << 3,
    GF_LIMITER = 1u << 4,
    GF_SHARED_SOURCE = 1u << 5,   // every lane reads the same source signal (broadcast loads)
    GF_WRITE_AUDIO = 1u << 6,
    GF_INPUT_DC_HP = 1u << 7,     // AF_INPUT_DC_HP80 applied inside the chain kernel
};

// Per-stream accumulators the finalize kernel reads (written once at the end of the render).
struct StreamAccum {
    double sum_in;
    double sum_out;
    float peak_in;
    float peak_out;
    float peak_pre_tp;
    float peak_out_tp;
    float limiter_gr_db;
    float tp_gr_db;
    float max_comp_gr;
    float max_de_gr;
    float effective_ceiling_db;
    uint32_t events;
    uint32_t non_finite;
    uint32_t n_rows;
    uint32_t n_samples;
    uint32_t reserved;
};

}  // namespace afsim
