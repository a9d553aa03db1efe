// afsim_kernels.cu -- sm_100a kernels of the batched chain simulator.
//
// chain_render_kernel: one WARP renders 32 streams (candidate x passage pairs), one stream per
// lane, start to end.  Time is cut into 32-sample tiles; inside a tile the chain runs STAGE BY
// STAGE over a shared-memory ring (`ring[row][lane]`, XOR-swizzled so that both the per-lane
// column walk of the recurrences and the transposed, coalesced global loads/stores of
// independent streams are bank-conflict free):
//
//   load -> [DC block + 80 Hz HP] -> de-esser <-> EQ (section-major) -> compressor
//        -> lookahead limiter (sliding max from per-tile maxima + one suffix scan per tile)
//        -> 4x true-peak limiter -> true-peak detector        (these three fused per 8 samples,
//           FIR windows and accumulators in registers, coefficients as FFMA immediates)
//
// Recurrence state is parked in lane-interleaved HBM/L2 tables between tiles (afsim_layout.h), so
// each stage's registers are live only while it runs.  Arithmetic follows the reference exactly:
// f64 state and coefficients, f32 hand-off between stages, no FMA contraction (-fmad=false) except
// the explicitly fused FIR (rust-core/src/dsp/true_peak.rs:173-186).
//
// finalize_kernel: one CTA per stream turns the per-block rows into the 30 metrics
// (python_api.rs:578-648): percentiles by bitonic sort in shared memory.
//
// Reference citations are relative to rust-core/src/.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/afsim.h"
#include "afsim_layout.h"

namespace afsim {

__constant__ float c_fir[4][32] = {
#include "true_peak_fir.inc"
};
// Constructor-default EQ coefficients (gain 0 at the default frequencies) the legacy band path
// fades out of (dsp/eq.rs:279-298 + dsp/biquad.rs:249-260); filled by the host per sample rate.
__constant__ double c_eq_default[10][5];

// ---------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ring_index(int row, int lane) { return (row << 5) + (lane ^ (row & 31)); }

__device__ __forceinline__ double lin_to_db(double v, double floor_) {  // dsp/util.rs:17-20
    return 20.0 * log10(fmax(fabs(v), floor_));
}
__device__ __forceinline__ double db_to_lin(double db) { return exp10(db / 20.0); }  // dsp/util.rs:11-14
__device__ __forceinline__ double clampd(double v, double lo, double hi) {
    return v < lo ? lo : (v > hi ? hi : v);
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ double smooth_ar(double prev, double in, double attack, double release) {
    const double c = in > prev ? attack : release;  // dsp/deesser.rs:147-154
    return c * prev + (1.0 - c) * in;
}
__device__ __forceinline__ double norm_range(double v, double s, double e) {
    return clampd((v - s) / (e - s), 0.0, 1.0);
}
__device__ __forceinline__ double lerpd(double a, double b, double t) { return a + (b - a) * t; }

struct Bq {  // one DF2T section, dsp/biquad.rs:262-274
    double b0, b1, b2, a1, a2;
};
__device__ __forceinline__ double bq_step(double x, const Bq& c, double& z1, double& z2) {
    const double y = c.b0 * x + z1;
    z1 = c.b1 * x - c.a1 * y + z2;
    z2 = c.b2 * x - c.a2 * y;
    return y;
}
__device__ __forceinline__ Bq load_bq(const double* __restrict__ tbl, int field, int lane) {
    Bq c;
    c.b0 = tbl[(field + 0) * kLanes + lane];
    c.b1 = tbl[(field + 1) * kLanes + lane];
    c.b2 = tbl[(field + 2) * kLanes + lane];
    c.a1 = tbl[(field + 3) * kLanes + lane];
    c.a2 = tbl[(field + 4) * kLanes + lane];
    return c;
}
__device__ __forceinline__ void store_bq(double* tbl, int field, int lane, const Bq& c) {
    tbl[(field + 0) * kLanes + lane] = c.b0;
    tbl[(field + 1) * kLanes + lane] = c.b1;
    tbl[(field + 2) * kLanes + lane] = c.b2;
    tbl[(field + 3) * kLanes + lane] = c.a1;
    tbl[(field + 4) * kLanes + lane] = c.a2;
}
#define PRM(field) params[(field) * kLanes + lane]
#define ST(field) state[(field) * kLanes + lane]
#define FST(field) fstate[(field) * kLanes + lane]

// A section in its first F samples: the constructor's filter fades into the configured one
// (dsp/biquad.rs:290-327).  `z` = {z1, z2, pending_z1, pending_z2}.
__device__ __forceinline__ double bq_step_fading(double x, const Bq& active, const Bq& pending, double* z, int n,
                                                 int fade_total) {
    const double ya = bq_step(x, active, z[0], z[1]);
    const double yp = bq_step(x, pending, z[2], z[3]);
    const double fade = (double)(n + 1) / (double)fade_total;
    return ya * (1.0 - fade) + yp * fade;
}

// ---------------------------------------------------------------------------------------------------
// stage: EQ cascade, section-major over the tile (dsp/eq.rs:317-322 is section-major too)
// ---------------------------------------------------------------------------------------------------
__device__ void stage_eq(float* ring, int ring_mask, const GroupHeader& hdr, const double* __restrict__ params,
                         double* state, int lane, int n0, int len) {
    const uint32_t lf = hdr.lane_flags[lane];
    const int nsec = (int)hdr.n_sections[lane];
    const int F = (int)hdr.fade_samples;
    const bool lane_fades = (lf & LF_EQ_FADE) != 0;
    const bool head = n0 < F && __any_sync(0xffffffffu, lane_fades);
    for (int s = 0; s < (int)hdr.max_sections; ++s) {
        if (s >= nsec) continue;  // lanes with fewer sections idle (no warp-level sync inside)
        const Bq tgt = load_bq(params, P_EQ + 5 * s, lane);
        double z[4];
        if (n0 == 0) {
            z[0] = z[1] = z[2] = z[3] = 0.0;
        } else {
            z[0] = ST(S_EQ + 4 * s + 0);
            z[1] = ST(S_EQ + 4 * s + 1);
            z[2] = ST(S_EQ + 4 * s + 2);
            z[3] = ST(S_EQ + 4 * s + 3);
        }
        if (!head) {
            double z1 = z[0], z2 = z[1];
#pragma unroll 4
            for (int t = 0; t < len; ++t) {
                const int idx = ring_index((n0 + t) & ring_mask, lane);
                const double x = (double)ring[idx];
                ring[idx] = (float)bq_step(x, tgt, z1, z2);
            }
            z[0] = z1;
            z[1] = z2;
        } else {
            Bq dflt;
            dflt.b0 = c_eq_default[s < 10 ? s : 0][0];
            dflt.b1 = c_eq_default[s < 10 ? s : 0][1];
            dflt.b2 = c_eq_default[s < 10 ? s : 0][2];
            dflt.a1 = c_eq_default[s < 10 ? s : 0][3];
            dflt.a2 = c_eq_default[s < 10 ? s : 0][4];
            for (int t = 0; t < len; ++t) {
                const int n = n0 + t;
                const int idx = ring_index(n & ring_mask, lane);
                const double x = (double)ring[idx];
                double y;
                if (lane_fades && n < F) {
                    y = bq_step_fading(x, dflt, tgt, z, n, F);
                    if (n + 1 == F) {  // promote_pending_coefficients
                        z[0] = z[2];
                        z[1] = z[3];
                    }
                } else {
                    y = bq_step(x, tgt, z[0], z[1]);
                }
                ring[idx] = (float)y;
            }
        }
        ST(S_EQ + 4 * s + 0) = z[0];
        ST(S_EQ + 4 * s + 1) = z[1];
        if (head) {
            ST(S_EQ + 4 * s + 2) = z[2];
            ST(S_EQ + 4 * s + 3) = z[3];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// stage: fixed input stage, DC block (f32) + 80 Hz high-pass (f64)  (audio/processor/routing.rs:826-843)
// ---------------------------------------------------------------------------------------------------
__device__ void stage_input_dc_hp(float* ring, int ring_mask, const double* __restrict__ params, double* state,
                                  float* fstate, int lane, int n0, int len) {
    const Bq hp = load_bq(params, P_IN_HP, lane);
    float x1 = 0.0f, y1 = 0.0f;
    double z1 = 0.0, z2 = 0.0;
    if (n0 != 0) {
        x1 = FST(FS_DC_X1);
        y1 = FST(FS_DC_Y1);
        z1 = ST(S_IN_HP + 0);
        z2 = ST(S_IN_HP + 1);
    }
    for (int t = 0; t < len; ++t) {
        const int idx = ring_index((n0 + t) & ring_mask, lane);
        const float in = ring[idx];
        const float out = __fadd_rn(__fsub_rn(in, x1), __fmul_rn(0.995f, y1));
        x1 = in;
        y1 = out;
        ring[idx] = (float)bq_step((double)out, hp, z1, z2);
    }
    FST(FS_DC_X1) = x1;
    FST(FS_DC_Y1) = y1;
    ST(S_IN_HP + 0) = z1;
    ST(S_IN_HP + 1) = z2;
}

// ---------------------------------------------------------------------------------------------------
// stage: dynamic-EQ de-esser (dsp/deesser.rs:405-547), sample-major
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ Bq design_peaking(double cs, double alpha, double gain_db) {  // dsp/biquad.rs:110-126,180-181
    const double a = exp10(gain_db / 40.0);
    const double b0 = 1.0 + alpha * a;
    const double b1 = -2.0 * cs;
    const double b2 = 1.0 - alpha * a;
    const double a0 = 1.0 + alpha / a;
    const double a1 = -2.0 * cs;
    const double a2 = 1.0 - alpha / a;
    Bq c;
    c.b0 = b0 / a0;
    c.b1 = b1 / a0;
    c.b2 = b2 / a0;
    c.a1 = a1 / a0;
    c.a2 = a2 / a0;
    return c;
}

__device__ __forceinline__ double de_confidence_target(double level_db, double voice_db, double narrowness) {
    // dsp/deesser.rs:171-219
    const double ratio_db = fmax(level_db - voice_db, 0.0);
    const double ratio_conf = norm_range(ratio_db, 1.5, 10.0);
    const double level_conf = norm_range(level_db, -62.0, -24.0);
    const double voice_conf = norm_range(voice_db, -58.0, -34.0);
    const double narrow_support = (ratio_db > 6.0 && level_db > -45.0) ? 0.75 : 0.0;
    const double voice_support = fmax(voice_conf, narrow_support);
    const double balance_conf = ratio_conf > 0.12 ? fmax(ratio_conf, voice_support * 0.65) : ratio_conf;
    const double broadband_penalty = lerpd(0.35, 1.0, balance_conf);
    const double narrowness_gain = lerpd(0.35, 1.0, norm_range(narrowness, 0.34, 0.68));
    return (0.62 * ratio_conf + 0.18 * level_conf + 0.20 * voice_support) * broadband_penalty * narrowness_gain;
}

__device__ void stage_deesser(float* ring, int ring_mask, const GroupHeader& hdr, const double* __restrict__ params,
                              double* state, float* rows_de, int lane, int n0, int len, int& blk, int& block_end) {
    const int F = (int)hdr.fade_samples;
    const int T = (int)hdr.n_samples;
    const bool auto_mode = (hdr.lane_flags[lane] & LF_DE_AUTO) != 0;
    const double det_attack = PRM(P_DE_DET_ATTACK), det_release = PRM(P_DE_DET_RELEASE);
    const double attack = PRM(P_DE_ATTACK), release = PRM(P_DE_RELEASE);
    const double max_red = PRM(P_DE_MAX_RED);

    // detector biquads: z[b][0..3] hp, z[b][4..7] lp ; dynamic EQ: dz[b][0..3]
    double dz[3][8], yz[3][4], env[3], conf[3], base[3], red[3], built_gain[3], cancel[3];
    Bq dyn[3];
    double broadband, current;
    if (n0 == 0) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dz[b][i] = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) yz[b][i] = 0.0;
            env[b] = conf[b] = base[b] = red[b] = built_gain[b] = cancel[b] = 0.0;
            dyn[b] = load_bq(params, P_DE_DYN0 + 5 * b, lane);
        }
        broadband = 0.0;
        current = 0.0;
    } else {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dz[b][i] = ST(S_DE_DET + 8 * b + i);
#pragma unroll
            for (int i = 0; i < 4; ++i) yz[b][i] = ST(S_DE_DYN + 4 * b + i);
            env[b] = ST(S_DE_ENV + b);
            conf[b] = ST(S_DE_CONF + b);
            base[b] = ST(S_DE_BASE + b);
            red[b] = ST(S_DE_RED + b);
            built_gain[b] = ST(S_DE_DYN_GAIN + b);
            cancel[b] = ST(S_DE_DYN_CANCEL + b);
            dyn[b] = load_bq(state, S_DE_DYN_COEF + 5 * b, lane);
        }
        broadband = ST(S_DE_BROADBAND);
        current = ST(S_DE_CURRENT);
    }
    const bool head = n0 < F;

    for (int t = 0; t < len; ++t) {
        const int n = n0 + t;
        const int idx = ring_index(n & ring_mask, lane);
        const float input = ring[idx];
        broadband = smooth_ar(broadband, (double)fabsf(input), det_attack, det_release);
        double level_db[3];
        double total_env = 0.0, max_env = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const Bq hp = load_bq(params, P_DE_DET + 10 * b, lane);
            const Bq lp = load_bq(params, P_DE_DET + 10 * b + 5, lane);
            float hp_out, sc;
            if (head && n < F) {
                const Bq hp0 = load_bq(params, P_DE_DET0 + 10 * b, lane);
                const Bq lp0 = load_bq(params, P_DE_DET0 + 10 * b + 5, lane);
                hp_out = (float)bq_step_fading((double)input, hp0, hp, &dz[b][0], n, F);
                sc = (float)bq_step_fading((double)hp_out, lp0, lp, &dz[b][4], n, F);
                if (n + 1 == F) {
                    dz[b][0] = dz[b][2];
                    dz[b][1] = dz[b][3];
                    dz[b][4] = dz[b][6];
                    dz[b][5] = dz[b][7];
                }
            } else {
                hp_out = (float)bq_step((double)input, hp, dz[b][0], dz[b][1]);
                sc = (float)bq_step((double)hp_out, lp, dz[b][4], dz[b][5]);
            }
            env[b] = smooth_ar(env[b], (double)fabsf(sc), det_attack, det_release);
            total_env += env[b];
            max_env = fmax(max_env, env[b]);
            level_db[b] = lin_to_db(env[b], 1e-10);
        }
        const double voice_level = fmax(broadband - total_env * 0.6, 1e-8);
        const double voice_db = lin_to_db(voice_level, 1e-10);
        const double narrowness = total_env > 1e-10 ? max_env / total_env : 0.0;

        double target[3];
        double target_sum = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double ratio_db = fmax(level_db[b] - voice_db, 0.0);
            const double dominance = max_env > 1e-10 ? sqrt(env[b] / max_env) : 0.0;
            const double conf_target = de_confidence_target(level_db[b], voice_db, narrowness) * dominance;
            conf[b] = smooth_ar(conf[b], clampd(conf_target, 0.0, 1.0), det_attack, det_release);
            double tr = 0.0;
            if (auto_mode) {
                const bool voice_active = voice_db > -55.0 || level_db[b] > -55.0;
                if (voice_active) {
                    const double base_target = clampd(ratio_db * 0.45, 0.0, 24.0);
                    const double c = base_target < base[b] ? PRM(P_DE_BASE_FALL) : PRM(P_DE_BASE_RISE);
                    base[b] = c * base[b] + (1.0 - c) * base_target;
                } else {
                    base[b] *= PRM(P_DE_BASE_INACTIVE);
                }
                const double conf_gain = norm_range(conf[b], PRM(P_DE_CONF_FLOOR), 1.0);
                const double over_db = fmax(ratio_db - base[b] - PRM(P_DE_TRIGGER), 0.0);
                tr = clampd(over_db * PRM(P_DE_SLOPE) * conf_gain, 0.0, PRM(P_DE_CAP));
            } else if (level_db[b] > PRM(P_DE_THRESHOLD)) {
                const double level_over = level_db[b] - PRM(P_DE_THRESHOLD);
                const double ratio_over = ratio_db - PRM(P_DE_RATIO_THR);
                if (ratio_over > 0.0) {
                    const double over_db = fmin(level_over, ratio_over);
                    const double conf_gain = norm_range(conf[b], 0.22, 1.0);
                    tr = clampd(PRM(P_DE_RATIO_FACTOR) * over_db * conf_gain, 0.0, max_red * 0.75);
                }
            }
            target[b] = tr;
            target_sum += tr;
        }
        if (target_sum > max_red && target_sum > 0.0) {
            const double scale = max_red / target_sum;
#pragma unroll
            for (int b = 0; b < 3; ++b) target[b] *= scale;
        }
        float processed = input;
        double total_red = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            red[b] = smooth_ar(red[b], target[b], attack, release);
            total_red += red[b];
            const double dyn_gain = -red[b];
            if (fabs(built_gain[b] - dyn_gain) > 0.001) {  // set_gain_db_immediate: cancels any fade, keeps z1/z2
                built_gain[b] = dyn_gain;
                dyn[b] = design_peaking(PRM(P_DE_DYN_COS + b), PRM(P_DE_DYN_ALPHA + b), dyn_gain);
                cancel[b] = 1.0;
            }
            double y;
            if (head && n < F && cancel[b] == 0.0) {
                const Bq pend = load_bq(params, P_DE_DYN1 + 5 * b, lane);
                y = bq_step_fading((double)processed, dyn[b], pend, yz[b], n, F);
                if (n + 1 == F) {
                    dyn[b] = pend;
                    yz[b][0] = yz[b][2];
                    yz[b][1] = yz[b][3];
                }
            } else {
                y = bq_step((double)processed, dyn[b], yz[b][0], yz[b][1]);
            }
            processed = (float)y;
        }
        current = fmin(total_red, max_red);
        ring[idx] = processed;
        if (n + 1 == block_end || n + 1 == T) {  // block-end sample of the meter (block_processor.rs:129-133)
            rows_de[blk] = (float)current;
            blk += 1;
            block_end += (int)hdr.block_samples;
        }
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) {
#pragma unroll
        for (int i = 0; i < 8; ++i) ST(S_DE_DET + 8 * b + i) = dz[b][i];
#pragma unroll
        for (int i = 0; i < 4; ++i) ST(S_DE_DYN + 4 * b + i) = yz[b][i];
        ST(S_DE_ENV + b) = env[b];
        ST(S_DE_CONF + b) = conf[b];
        ST(S_DE_BASE + b) = base[b];
        ST(S_DE_RED + b) = red[b];
        ST(S_DE_DYN_GAIN + b) = built_gain[b];
        ST(S_DE_DYN_CANCEL + b) = cancel[b];
        store_bq(state, S_DE_DYN_COEF + 5 * b, lane, dyn[b]);
    }
    ST(S_DE_BROADBAND) = broadband;
    ST(S_DE_CURRENT) = current;
}

// ---------------------------------------------------------------------------------------------------
// stage: compressor (dsp/compressor.rs:725-774), sample-major
// ---------------------------------------------------------------------------------------------------
__device__ void stage_compressor(float* ring, int ring_mask, const GroupHeader& hdr, const double* __restrict__ params,
                                 double* state, float* rows_comp, int lane, int n0, int len, int& blk,
                                 int& block_end) {
    const int T = (int)hdr.n_samples;
    const uint32_t lf = hdr.lane_flags[lane];
    const bool adaptive = (lf & LF_C_ADAPTIVE) != 0;
    const bool sidechain = (lf & LF_C_SIDECHAIN) != 0;
    const double threshold = PRM(P_C_THRESHOLD), factor = PRM(P_C_FACTOR), knee = PRM(P_C_KNEE);
    const double attack = PRM(P_C_ATTACK), det_release = PRM(P_C_DET_RELEASE), release = PRM(P_C_RELEASE);
    const double rms_c = PRM(P_C_RMS), makeup_lin = PRM(P_C_MAKEUP_LIN), sc_c = PRM(P_C_SC_COEFF);
    const double band_c = PRM(P_C_BAND), fast_c = PRM(P_C_FAST), charge_c = PRM(P_C_CHARGE), slow_c = PRM(P_C_SLOW);
    const double knee_half = knee / 2.0;
    const double knee_start = threshold - knee_half, knee_end = threshold + knee_half;

    double prev_in = 0.0, prev_out = 0.0, low_sq = 0.0, voiced_sq = 0.0, presence_sq = 0.0;
    double peak_env = -120.0, rms_env = 0.0, gr = 0.0, fast_env = 0.0, slow_env = 0.0;
    if (n0 != 0) {
        prev_in = ST(S_C_PREV_IN);
        prev_out = ST(S_C_PREV_OUT);
        low_sq = ST(S_C_LOW);
        voiced_sq = ST(S_C_VOICED);
        presence_sq = ST(S_C_PRESENCE);
        peak_env = ST(S_C_PEAK_ENV);
        rms_env = ST(S_C_RMS_ENV);
        gr = ST(S_C_GR);
        fast_env = ST(S_C_FAST);
        slow_env = ST(S_C_SLOW);
    }
    for (int t = 0; t < len; ++t) {
        const int n = n0 + t;
        const int idx = ring_index(n & ring_mask, lane);
        const double x = (double)ring[idx];
        double det = x, weight_db = 0.0;
        if (sidechain) {  // :407-450
            det = sc_c * (prev_out + x - prev_in);
            prev_in = x;
            prev_out = det;
            const double low = x - det;
            const double presence = 0.65 * det + 0.35 * (det - low);
            low_sq = band_c * low_sq + (1.0 - band_c) * low * low;
            voiced_sq = band_c * voiced_sq + (1.0 - band_c) * det * det;
            presence_sq = band_c * presence_sq + (1.0 - band_c) * presence * presence;
            const double low_rms = sqrt(low_sq);
            const double voiced_rms = fmax(sqrt(voiced_sq), 1e-8);
            const double presence_rms = sqrt(presence_sq);
            const double plosive = clampd(low_rms / voiced_rms, 0.0, 32.0);
            const double amount = clampd((plosive - 1.25) / (5.0 - 1.25), 0.0, 1.0);
            const double penalty = 1.0 - amount * (1.0 - 0.35);
            const double presence_ratio = clampd(presence_rms / voiced_rms, 0.0, 4.0);
            const double pw = 1.0 + 0.18 * clampd(presence_ratio - 0.75, 0.0, 1.0);
            const double weight = clampd(penalty * pw, 0.35, 1.15);
            weight_db = lin_to_db(weight, 1e-10);
        }
        const double inst_peak_db = lin_to_db(fabs(det), 1e-10);
        const double pc = inst_peak_db > peak_env ? attack : det_release;
        peak_env = pc * peak_env + (1.0 - pc) * inst_peak_db;
        rms_env = rms_c * rms_env + (1.0 - rms_c) * (det * det);
        const double rms_db = lin_to_db(sqrt(rms_env), 1e-10);
        const double blended = 0.6 * db_to_lin(peak_env) + 0.4 * db_to_lin(rms_db);  // :681-686
        const double detector_db = lin_to_db(blended, 1e-10) + weight_db;

        double target;  // compute_gain_reduction :657-678
        if (knee <= 0.0) {
            target = detector_db <= threshold ? 0.0 : (detector_db - threshold) * factor;
        } else if (detector_db <= knee_start) {
            target = 0.0;
        } else if (detector_db >= knee_end) {
            target = (detector_db - threshold) * factor;
        } else {
            const double k = detector_db - knee_start;
            target = factor * k * k / (2.0 * knee);
        }
        if (!adaptive) {  // smooth_gain_reduction :468-505
            const double c = target > gr ? attack : release;
            gr = c * gr + (1.0 - c) * target;
        } else {
            if (target > gr)
                fast_env = attack * gr + (1.0 - attack) * target;
            else
                fast_env = fast_c * fast_env + (1.0 - fast_c) * target;
            if (target > 3.0)
                slow_env = charge_c * slow_env + (1.0 - charge_c) * target;
            else
                slow_env *= slow_c;
            gr = fmax(fast_env, slow_env);
        }
        const double gain = db_to_lin(-gr) * makeup_lin;
        ring[idx] = (float)(x * gain);
        if (n + 1 == block_end || n + 1 == T) {
            rows_comp[blk] = (float)gr;
            blk += 1;
            block_end += (int)hdr.block_samples;
        }
    }
    ST(S_C_PREV_IN) = prev_in;
    ST(S_C_PREV_OUT) = prev_out;
    ST(S_C_LOW) = low_sq;
    ST(S_C_VOICED) = voiced_sq;
    ST(S_C_PRESENCE) = presence_sq;
    ST(S_C_PEAK_ENV) = peak_env;
    ST(S_C_RMS_ENV) = rms_env;
    ST(S_C_GR) = gr;
    ST(S_C_FAST) = fast_env;
    ST(S_C_SLOW) = slow_env;
}

// ---------------------------------------------------------------------------------------------------
// 4x polyphase FIR over a register window: 8 outputs x 4 phases, taps accumulated k = 0..31 in the
// reference's order with fused multiply-add (dsp/true_peak.rs:173-186).  w[i] = x[n0 - 31 + i].
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fir8_peaks(const float (&w)[39], float (&peak)[8]) {
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[j][p] = 0.0f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float c = c_fir[p][k];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j][p] = __fmaf_rn(c, w[31 + j - k], acc[j][p]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float pk = fabsf(w[31 + j]);
#pragma unroll
        for (int p = 0; p < 4; ++p) pk = fmaxf(pk, fabsf(acc[j][p]));
        peak[j] = pk;
    }
}

// ---------------------------------------------------------------------------------------------------
// stage: lookahead limiter -> true-peak limiter -> true-peak detector + output statistics
// (dsp/limiter.rs:246-284, dsp/true_peak.rs:337-378,208-218, block_processor.rs:146-159,
//  python_api.rs:529-575)
// ---------------------------------------------------------------------------------------------------
struct OutputCtx {
    float* rows_out;      // row.1 (output RMS dB) of this lane's stream
    float* audio;         // lane's output audio or nullptr
    float* stage_tile;    // [32][32] swizzled scratch (suffix maxima, then transposed audio staging)
    float* tile_max;      // [kTileMaxSlots][32] per-tile |x| maxima of the limiter input
};

__device__ void stage_output(float* ring, int ring_mask, const GroupHeader& hdr, const double* __restrict__ params,
                             const float* __restrict__ fparams, double* state, float* fstate, const OutputCtx& ctx,
                             int lane, int n0, int len, int& blk, int& block_end) {
    const int T = (int)hdr.n_samples;
    const int L = (int)hdr.lookahead;
    const bool limiter_on = (hdr.flags & GF_LIMITER) != 0;
    const double ceil_lin = PRM(P_L_CEIL), rel = PRM(P_L_RELEASE);
    const float tp_ceil = fparams[FP_TP_CEIL * kLanes + lane];
    const float tp_rel = fparams[FP_TP_RELEASE * kLanes + lane];

    float win_in[39], win_out[39];
    double g_lim = 1.0, min_g_lim = 1.0, sum_out = 0.0, blk_out = 0.0;
    float g_tp = 1.0f, min_g_tp = 1.0f, limited = 0.0f, peak_out = 0.0f, peak_pre = 0.0f, peak_tp = 0.0f;
    float events = 0.0f, nonfinite = 0.0f;
    if (n0 == 0) {
#pragma unroll
        for (int i = 0; i < 31; ++i) {
            win_in[i] = 0.0f;
            win_out[i] = 0.0f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 31; ++i) {
            win_in[i] = FST(FS_WIN_IN + i);
            win_out[i] = FST(FS_WIN_OUT + i);
        }
        g_lim = ST(S_L_GAIN);
        min_g_lim = ST(S_L_MIN_GAIN);
        sum_out = ST(S_SUM_OUT);
        blk_out = ST(S_BLK_OUT);
        g_tp = FST(FS_TP_GAIN);
        min_g_tp = FST(FS_TP_MIN_GAIN);
        limited = FST(FS_TP_BLOCK_LIMITED);
        peak_out = FST(FS_PEAK_OUT);
        peak_pre = FST(FS_PEAK_PRE_TP);
        peak_tp = FST(FS_PEAK_OUT_TP);
        events = FST(FS_EVENTS);
        nonfinite = FST(FS_NONFINITE);
    }

    // ---- sliding-window maximum of |x| over [n-L, n] from per-tile maxima ---------------------------
    // Tiles are 32-aligned.  For L >= 32 the window start a = n-L lives in an older tile: window max =
    // max(suffix max of a's tile from a, maxima of the whole tiles in between, running prefix max of
    // the current tile).  a's tile changes at most once inside the tile -> two segments.
    const int k_cur = n0 >> 5;
    float prefix_max = 0.0f;  // running max of |x| over the current tile so far

    for (int g0 = 0; g0 < len; g0 += 8) {
        float lim_out[8];
        if (limiter_on) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = g0 + j;
                const int n = n0 + t;
                float out_v = 0.0f;
                if (t < len) {
                    const float in = ring[ring_index(n & ring_mask, lane)];
                    const float in_abs = fabsf(in);
                    float window = 0.0f;
                    if (L >= 32) {
                        const int a = n - L;                       // may be negative: zeros
                        const int ka = a >> 5;                     // floor division
                        if (a >= 0) window = ctx.stage_tile[ring_index(((ka & 1) << 5 | (a & 31)) & 63, lane)];
                        // whole tiles strictly between ka and k_cur
                        for (int k = (ka < -1 ? -1 : ka) + 1; k < k_cur; ++k)
                            window = fmaxf(window, ctx.tile_max[(k % kTileMaxSlots) * kLanes + lane]);
                        window = fmaxf(window, prefix_max);
                    } else {
                        for (int d = 1; d <= L; ++d) {
                            const int m = n - d;
                            if (m >= 0) window = fmaxf(window, fabsf(ring[ring_index(m & ring_mask, lane)]));
                        }
                    }
                    prefix_max = fmaxf(prefix_max, in_abs);
                    const double peak = fmax((double)window, (double)in_abs);
                    const double target = peak > ceil_lin ? ceil_lin / peak : 1.0;
                    if (target < g_lim)
                        g_lim = target;
                    else
                        g_lim = rel * g_lim + (1.0 - rel) * target;
                    min_g_lim = fmin(min_g_lim, g_lim);
                    const int md = n - L;
                    const double delayed = md >= 0 ? (double)ring[ring_index(md & ring_mask, lane)] : 0.0;
                    out_v = (float)clampd(delayed * g_lim, -ceil_lin, ceil_lin);
                }
                lim_out[j] = out_v;
            }
            // ---- true-peak limiter (f32) ---------------------------------------------------------------
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float v = lim_out[j];
                win_in[31 + j] = isfinite(v) ? v : 0.0f;
            }
            float itp[8];
            fir8_peaks(win_in, itp);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = g0 + j;
                float o = 0.0f;
                if (t < len) {
                    peak_pre = fmaxf(peak_pre, itp[j]);
                    const float target =
                        itp[j] > tp_ceil ? clampf(__fdiv_rn(__fmul_rn(tp_ceil, 0.999f), itp[j]), 0.0f, 1.0f) : 1.0f;
                    if (target < g_tp) {
                        g_tp = target;
                        limited = 1.0f;
                    } else {
                        g_tp = __fadd_rn(__fmul_rn(tp_rel, g_tp), __fmul_rn(__fsub_rn(1.0f, tp_rel), target));
                    }
                    min_g_tp = fminf(min_g_tp, g_tp);
                    o = clampf(__fmul_rn(win_in[11 + j], g_tp), -tp_ceil, tp_ceil);  // delay 20 = window slot 31+j-20
                    if (!isfinite(o)) o = 0.0f;
                }
                lim_out[j] = o;
            }
#pragma unroll
            for (int i = 0; i < 31; ++i) win_in[i] = win_in[i + 8];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int t = g0 + j;
                lim_out[j] = t < len ? ring[ring_index((n0 + t) & ring_mask, lane)] : 0.0f;
            }
        }
        // ---- output statistics + true-peak detector ---------------------------------------------------
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float v = lim_out[j];
            win_out[31 + j] = isfinite(v) ? v : 0.0f;
        }
        float otp[8];
        fir8_peaks(win_out, otp);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int t = g0 + j;
            if (t < len) {
                const int n = n0 + t;
                const float v = lim_out[j];
                peak_out = fmaxf(peak_out, fabsf(v));
                peak_tp = fmaxf(peak_tp, otp[j]);
                const double sq = (double)v * (double)v;
                sum_out += sq;
                if (isfinite(v))
                    blk_out += sq;
                else
                    nonfinite = 1.0f;
                if (ctx.audio) ctx.stage_tile[ring_index(64 + t, lane)] = v;
                if (n + 1 == block_end || n + 1 == T) {
                    const int blen = n + 1 - (block_end - (int)hdr.block_samples);
                    const float rms = (float)sqrt(blk_out / (double)blen);
                    ctx.rows_out[blk] = __fmul_rn(20.0f, log10f(fmaxf(rms, 1.0e-12f)));
                    blk_out = 0.0;
                    events += limited;
                    limited = 0.0f;
                    blk += 1;
                    block_end += (int)hdr.block_samples;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 31; ++i) win_out[i] = win_out[i + 8];
    }
    if (limiter_on && L >= 32) ctx.tile_max[(k_cur % kTileMaxSlots) * kLanes + lane] = prefix_max;

#pragma unroll
    for (int i = 0; i < 31; ++i) {
        FST(FS_WIN_IN + i) = win_in[i];
        FST(FS_WIN_OUT + i) = win_out[i];
    }
    ST(S_L_GAIN) = g_lim;
    ST(S_L_MIN_GAIN) = min_g_lim;
    ST(S_SUM_OUT) = sum_out;
    ST(S_BLK_OUT) = blk_out;
    FST(FS_TP_GAIN) = g_tp;
    FST(FS_TP_MIN_GAIN) = min_g_tp;
    FST(FS_TP_BLOCK_LIMITED) = limited;
    FST(FS_PEAK_OUT) = peak_out;
    FST(FS_PEAK_PRE_TP) = peak_pre;
    FST(FS_PEAK_OUT_TP) = peak_tp;
    FST(FS_EVENTS) = events;
    FST(FS_NONFINITE) = nonfinite;
}

// Suffix maxima of |x| for the (at most two) tiles the limiter's window start walks through while the
// warp processes tile k_cur: slot (k & 1) of the 64-row scratch holds suffix maxima of tile k.
__device__ void limiter_prepare_suffix(const float* ring, int ring_mask, float* stage_tile, int lane, int n0, int L) {
    const int a0 = n0 - L;
    const int ka0 = a0 >> 5;
    for (int seg = 0; seg < 2; ++seg) {
        const int k = ka0 + seg;
        if (k < 0) continue;
        if (seg == 1 && (a0 & 31) == 0) continue;  // aligned: the window start never leaves tile ka0
        float m = 0.0f;
        for (int i = 31; i >= 0; --i) {
            m = fmaxf(m, fabsf(ring[ring_index(((k << 5) + i) & ring_mask, lane)]));
            stage_tile[ring_index(((k & 1) << 5) | i, lane)] = m;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// the render kernel
// ---------------------------------------------------------------------------------------------------
struct RenderArgs {
    const GroupHeader* headers;
    const double* params;   // [group][P_COUNT][32]
    const float* fparams;   // [group][FP_COUNT][32]
    double* state;          // [group][S_COUNT][32]
    float* fstate;          // [group][FS_COUNT][32]
    const float* signals;   // source signal pool
    float* audio;           // output audio pool (or nullptr)
    float* rows;            // [stream][4][rows_pitch]
    StreamAccum* accum;     // [stream]
    int n_groups;
    int ring_rows;          // power of two >= lookahead + 64
};

extern __shared__ float smem_dyn[];

__global__ void __launch_bounds__(32) chain_render_kernel(RenderArgs args) {
    const int group = blockIdx.x;
    if (group >= args.n_groups) return;
    const int lane = threadIdx.x;
    const GroupHeader& hdr = args.headers[group];
    const double* params = args.params + (size_t)group * P_COUNT * kLanes;
    const float* fparams = args.fparams + (size_t)group * FP_COUNT * kLanes;
    double* state = args.state + (size_t)group * S_COUNT * kLanes;
    float* fstate = args.fstate + (size_t)group * FS_COUNT * kLanes;

    float* ring = smem_dyn;                                   // [ring_rows][32]
    float* stage_tile = ring + (size_t)args.ring_rows * kLanes;  // [96][32]: 2 suffix tiles + audio staging
    float* tile_max = stage_tile + 96 * kLanes;                // [kTileMaxSlots][32]
    const int ring_mask = args.ring_rows - 1;
    for (int r = 0; r < args.ring_rows; ++r) ring[(r << 5) + lane] = 0.0f;
    for (int r = 0; r < kTileMaxSlots; ++r) tile_max[(r << 5) + lane] = 0.0f;
    for (int r = 0; r < 96; ++r) stage_tile[(r << 5) + lane] = 0.0f;
    __syncwarp();

    const int T = (int)hdr.n_samples;
    const uint32_t gflags = hdr.flags;
    const bool active = (hdr.lane_flags[lane] & LF_ACTIVE) != 0;
    const uint32_t slot = hdr.first_stream + (uint32_t)lane;
    float* rows = args.rows + (size_t)slot * 4 * hdr.rows_pitch;
    float* rows_in = rows;  // padding lanes own scratch slots, so every lane may write its rows
    const float* src = args.signals + hdr.src_offset[lane];
    const uint64_t src_off = hdr.src_offset[lane];
    float* audio = (gflags & GF_WRITE_AUDIO) && active ? args.audio + hdr.audio_offset[lane] : nullptr;
    const uint64_t audio_off = hdr.audio_offset[lane];
    double sum_in = 0.0, blk_in = 0.0;
    float peak_in = 0.0f;
    int blk_load = 0, end_load = (int)hdr.block_samples;
    int blk_de = 0, end_de = (int)hdr.block_samples;
    int blk_c = 0, end_c = (int)hdr.block_samples;
    int blk_o = 0, end_o = (int)hdr.block_samples;
    // Padding lanes (LF_ACTIVE clear) replicate a real lane's parameters and write to scratch slots.

    for (int n0 = 0; n0 < T; n0 += kTile) {
        const int len = min(kTile, T - n0);
        // ---- load + sanitize (python_api.rs:517-524) ------------------------------------------------
        if (gflags & GF_SHARED_SOURCE) {
            const float mine = lane < len ? __ldg(src + n0 + lane) : 0.0f;
            for (int t = 0; t < len; ++t) {
                float v = __shfl_sync(0xffffffffu, mine, t);
                if (!isfinite(v)) v = 0.0f;
                ring[ring_index((n0 + t) & ring_mask, lane)] = v;
            }
        } else {
            // transposed, coalesced: for each stream j of the warp, lanes read 32 consecutive samples
            for (int j = 0; j < kLanes; ++j) {
                const uint64_t off = __shfl_sync(0xffffffffu, src_off, j);
                if (lane < len) {
                    float v = __ldg(args.signals + off + n0 + lane);
                    if (!isfinite(v)) v = 0.0f;
                    ring[ring_index((n0 + lane) & ring_mask, j)] = v;
                }
            }
        }
        __syncwarp();
        if (gflags & GF_INPUT_DC_HP) stage_input_dc_hp(ring, ring_mask, params, state, fstate, lane, n0, len);
        for (int t = 0; t < len; ++t) {
            const int n = n0 + t;
            const float v = ring[ring_index(n & ring_mask, lane)];
            const double sq = (double)v * (double)v;
            sum_in += sq;
            blk_in += sq;
            peak_in = fmaxf(peak_in, fabsf(v));
            if (n + 1 == end_load || n + 1 == T) {
                const int blen = n + 1 - (end_load - (int)hdr.block_samples);
                const float rms = (float)sqrt(blk_in / (double)blen);
                rows_in[blk_load] = __fmul_rn(20.0f, log10f(fmaxf(rms, 1.0e-12f)));
                blk_in = 0.0;
                blk_load += 1;
                end_load += (int)hdr.block_samples;
            }
        }
        // ---- de-esser <-> EQ (block_processor.rs:125-141) ---------------------------------------------
        if (gflags & GF_EQ_BEFORE_DEESSER) {
            if (gflags & GF_EQ) stage_eq(ring, ring_mask, hdr, params, state, lane, n0, len);
            if (gflags & GF_DEESSER)
                stage_deesser(ring, ring_mask, hdr, params, state, rows + 3 * hdr.rows_pitch, lane, n0, len, blk_de, end_de);
        } else {
            if (gflags & GF_DEESSER)
                stage_deesser(ring, ring_mask, hdr, params, state, rows + 3 * hdr.rows_pitch, lane, n0, len, blk_de, end_de);
            if (gflags & GF_EQ) stage_eq(ring, ring_mask, hdr, params, state, lane, n0, len);
        }
        if (gflags & GF_COMPRESSOR)
            stage_compressor(ring, ring_mask, hdr, params, state, rows + 2 * hdr.rows_pitch, lane, n0, len, blk_c, end_c);
        // ---- limiter + true peak ------------------------------------------------------------------------
        if ((gflags & GF_LIMITER) && (int)hdr.lookahead >= 32)
            limiter_prepare_suffix(ring, ring_mask, stage_tile, lane, n0, (int)hdr.lookahead);
        OutputCtx ctx;
        ctx.rows_out = rows + 1 * hdr.rows_pitch;
        ctx.audio = (gflags & GF_WRITE_AUDIO) ? (float*)1 : nullptr;
        ctx.stage_tile = stage_tile;
        ctx.tile_max = tile_max;
        stage_output(ring, ring_mask, hdr, params, fparams, state, fstate, ctx, lane, n0, len, blk_o, end_o);
        if (gflags & GF_WRITE_AUDIO) {
            __syncwarp();
            for (int j = 0; j < kLanes; ++j) {
                const uint64_t off = __shfl_sync(0xffffffffu, audio_off, j);
                const uint32_t lf_j = __shfl_sync(0xffffffffu, hdr.lane_flags[lane], j);
                if ((lf_j & LF_ACTIVE) && lane < len)
                    args.audio[off + n0 + lane] = stage_tile[ring_index(64 + lane, j)];
            }
            __syncwarp();
        }
    }
    (void)audio;

    // ---- per-stream accumulators -> finalize kernel -----------------------------------------------------
    {
        StreamAccum acc;
        acc.sum_in = sum_in;
        acc.peak_in = peak_in;
        if (T > 0) {
            acc.sum_out = ST(S_SUM_OUT);
            acc.peak_out = FST(FS_PEAK_OUT);
            acc.peak_pre_tp = FST(FS_PEAK_PRE_TP);
            acc.peak_out_tp = FST(FS_PEAK_OUT_TP);
            const double min_g = ST(S_L_MIN_GAIN);
            const float min_g_tp = FST(FS_TP_MIN_GAIN);
            const bool lim = (gflags & GF_LIMITER) != 0;
            // dsp/limiter.rs:272-279 and dsp/true_peak.rs:320-326: the per-sample reduction is monotone in
            // the gain, so the block maxima collapse to one conversion of the minimum gain.
            acc.limiter_gr_db = (lim && min_g < 1.0) ? (float)(-lin_to_db(min_g, 1e-10)) : 0.0f;
            acc.tp_gr_db = (lim && min_g_tp < 1.0f) ? __fmul_rn(-20.0f, log10f(fmaxf(min_g_tp, 1e-10f))) : 0.0f;
            acc.events = (uint32_t)FST(FS_EVENTS);
            acc.non_finite = FST(FS_NONFINITE) != 0.0f ? 1u : 0u;
        } else {
            acc.sum_out = 0.0;
            acc.peak_out = acc.peak_pre_tp = acc.peak_out_tp = 0.0f;
            acc.limiter_gr_db = acc.tp_gr_db = 0.0f;
            acc.events = 0;
            acc.non_finite = 0;
        }
        acc.max_comp_gr = 0.0f;  // maxima of the block-end rows are taken in finalize
        acc.max_de_gr = 0.0f;
        acc.effective_ceiling_db = 0.0f;
        acc.n_rows = (uint32_t)((T + (int)hdr.block_samples - 1) / (int)hdr.block_samples);
        acc.n_samples = (uint32_t)T;
        acc.reserved = 0;
        args.accum[slot] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------
// finalize: rows -> metrics (python_api.rs:578-713)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int total_key(float v) {  // f32::total_cmp order as a signed int
    int i = __float_as_int(v);
    return i ^ (int)(((unsigned)(i >> 31)) >> 1);
}

// Block-wide bitonic sort of n floats (padded with +NaN-max keys) by total_cmp; buf has n_pad entries.
__device__ void block_sort(float* buf, int n, int n_pad) {
    for (int i = threadIdx.x + n; i < n_pad; i += blockDim.x) buf[i] = __int_as_float(0x7fffffff);
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float a = buf[i], b = buf[ixj];
                    const bool up = (i & k) == 0;
                    const bool gt = total_key(a) > total_key(b);
                    if (gt == up) {
                        buf[i] = b;
                        buf[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// percentile_f32 (python_api.rs:58-72) of an already sorted buffer.
__device__ float sorted_percentile(const float* buf, int n, float p) {
    if (n == 0) return 0.0f;
    const float position = __fmul_rn((float)(n - 1), clampf(p, 0.0f, 1.0f));
    const int lower = (int)floorf(position);
    const int upper = (int)ceilf(position);
    if (lower == upper) return buf[lower];
    const float fraction = __fsub_rn(position, (float)lower);
    return __fadd_rn(buf[lower], __fmul_rn(fraction, __fsub_rn(buf[upper], buf[lower])));
}

struct FinalizeArgs {
    const float* rows;            // [stream][4][rows_pitch]
    const StreamAccum* accum;
    const float* ceilings;        // effective ceiling dB per stream (f32, python_api.rs:472-473)
    AfChainMetrics* metrics;
    float* scratch;               // [stream][2][n_pad] when rows do not fit shared memory
    int rows_pitch;
    int n_pad;                    // power of two >= max rows
    int use_global_scratch;
    int n_streams;
};

extern __shared__ float fin_smem[];

__device__ int compact_into(float* dst, int n_rows, const float* rows_in, const float* rows_val0,
                            const float* rows_val1, int mode, float thr, int* counter) {
    // Order-preserving compaction is not needed (everything is sorted next), but a deterministic
    // result is: one thread walks the rows.  n_rows <= a few thousand.
    __syncthreads();
    if (threadIdx.x == 0) {
        int c = 0;
        for (int i = 0; i < n_rows; ++i) {
            const float in_db = rows_in[i];
            bool keep;
            float v;
            switch (mode) {
                case 0: keep = in_db >= thr; v = fmaxf(rows_val0[i], 0.0f); break;                 // active GR
                case 1: keep = true; v = fmaxf(rows_val0[i], 0.0f); break;                          // all GR
                case 2: keep = in_db >= thr && in_db > -100.0f; v = __fsub_rn(rows_val1[i], in_db); break;  // active gain
                case 3: keep = in_db < thr && in_db > -100.0f; v = __fsub_rn(rows_val1[i], in_db); break;   // silence delta
                default: keep = in_db < thr; v = -fmaxf(rows_val0[i], 0.0f); break;                 // silence gain
            }
            if (keep) dst[c++] = v;
        }
        *counter = c;
    }
    __syncthreads();
    return *counter;
}

__global__ void __launch_bounds__(256) finalize_kernel(FinalizeArgs args) {
    const int s = blockIdx.x;
    if (s >= args.n_streams) return;
    __shared__ int counter;
    __shared__ float sh[16];
    const StreamAccum acc = args.accum[s];
    const int n_rows = (int)acc.n_rows;
    const float* r_in = args.rows + (size_t)s * 4 * args.rows_pitch;
    const float* r_out = r_in + args.rows_pitch;
    const float* r_comp = r_out + args.rows_pitch;
    const float* r_de = r_comp + args.rows_pitch;
    float* buf = args.use_global_scratch ? args.scratch + (size_t)s * 2 * args.n_pad : fin_smem;
    float* buf2 = buf + args.n_pad;
    int n_pad = 1;
    while (n_pad < n_rows) n_pad <<= 1;
    if (n_pad < 2) n_pad = 2;

    // input RMS rows -> p20 / p90 -> active threshold
    for (int i = threadIdx.x; i < n_rows; i += blockDim.x) buf[i] = r_in[i];
    __syncthreads();
    block_sort(buf, n_rows, n_pad);
    if (threadIdx.x == 0) {
        const float floor_db = sorted_percentile(buf, n_rows, 0.20f);
        const float p90 = sorted_percentile(buf, n_rows, 0.90f);
        sh[0] = fmaxf(fmaxf(__fadd_rn(floor_db, 6.0f), __fsub_rn(p90, 24.0f)), -60.0f);
    }
    __syncthreads();
    const float thr = sh[0];

    // active compressor / de-esser reductions
    int n_act = compact_into(buf, n_rows, r_in, r_comp, nullptr, 0, thr, &counter);
    int mode = 0;
    if (n_act < 3) {
        mode = 1;
        n_act = compact_into(buf, n_rows, r_in, r_comp, nullptr, 1, thr, &counter);
    }
    if (threadIdx.x == 0) {
        int c = 0;
        for (int i = 0; i < n_act; ++i) c += buf[i] >= 0.10f ? 1 : 0;
        sh[1] = n_act > 0 ? __fdiv_rn((float)c, (float)n_act) : 0.0f;
    }
    int np = 2;
    while (np < n_act) np <<= 1;
    block_sort(buf, n_act, np);
    if (threadIdx.x == 0) {
        sh[2] = sorted_percentile(buf, n_act, 0.50f);
        sh[3] = sorted_percentile(buf, n_act, 0.95f);
    }
    __syncthreads();
    int n_de = compact_into(buf, n_rows, r_in, r_de, nullptr, mode, thr, &counter);
    np = 2;
    while (np < n_de) np <<= 1;
    block_sort(buf, n_de, np);
    if (threadIdx.x == 0) {
        sh[4] = sorted_percentile(buf, n_de, 0.50f);
        sh[5] = sorted_percentile(buf, n_de, 0.95f);
    }
    __syncthreads();
    // active output gain, silence level delta, silence output gain: medians
    for (int q = 0; q < 3; ++q) {
        const int m = q == 0 ? 2 : (q == 1 ? 3 : 4);
        const int cnt = compact_into(buf, n_rows, r_in, r_comp, r_out, m, thr, &counter);
        np = 2;
        while (np < cnt) np <<= 1;
        block_sort(buf, cnt, np);
        if (threadIdx.x == 0) sh[6 + q] = sorted_percentile(buf, cnt, 0.50f);
        __syncthreads();
    }
    // compressor pumping score (python_api.rs:74-111) over all rows' clamped GR at 50 Hz
    float pumping = 0.0f;
    if (n_rows >= 3) {
        if (threadIdx.x == 0) {
            const float dt = __fdiv_rn(1.0f, 50.0f);
            const float two_pi = __fmul_rn(2.0f, 3.14159265358979323846f);
            const float hp_rc = __fdiv_rn(1.0f, __fmul_rn(two_pi, 2.0f));
            const float lp_rc = __fdiv_rn(1.0f, __fmul_rn(two_pi, 8.0f));
            const float hp_alpha = __fdiv_rn(hp_rc, __fadd_rn(hp_rc, dt));
            const float lp_alpha = __fdiv_rn(dt, __fadd_rn(lp_rc, dt));
            float prev = fmaxf(r_comp[0], 0.0f), hp = 0.0f, bp = 0.0f;
            int bad = 0;
            for (int i = 1; i < n_rows; ++i) {
                const float v = fmaxf(r_comp[i], 0.0f);
                if (!isfinite(v)) {
                    bad = 1;
                    break;
                }
                hp = __fmul_rn(hp_alpha, __fsub_rn(__fadd_rn(hp, v), prev));
                bp = __fadd_rn(bp, __fmul_rn(lp_alpha, __fsub_rn(hp, bp)));
                buf[i - 1] = fabsf(bp);
                buf2[i - 1] = fabsf(__fsub_rn(v, prev));
                prev = v;
            }
            sh[10] = (float)bad;
        }
        __syncthreads();
        if (sh[10] != 0.0f) {
            pumping = INFINITY;
        } else {
            const int m = n_rows - 1;
            // robust RMS needs the unsorted band-pass trace after its p95 is known: sum first with the
            // limit from a sorted copy.  Keep the unsorted trace in global rows? -> sort buf2-copy instead.
            // Step 1: p95 of deltas (buf2), step 2: copy buf -> buf2, sort buf2 for the limit.
            np = 2;
            while (np < m) np <<= 1;
            block_sort(buf2, m, np);
            if (threadIdx.x == 0) sh[11] = sorted_percentile(buf2, m, 0.95f);
            __syncthreads();
            for (int i = threadIdx.x; i < m; i += blockDim.x) buf2[i] = buf[i];
            __syncthreads();
            block_sort(buf2, m, np);
            if (threadIdx.x == 0) {
                const float limit = sorted_percentile(buf2, m, 0.95f);
                float sum = 0.0f;
                for (int i = 0; i < m; ++i) {
                    const float v = fminf(buf[i], limit);
                    sum = __fadd_rn(sum, __fmul_rn(v, v));
                }
                sh[12] = __fadd_rn(sqrtf(__fdiv_rn(sum, (float)m)), sh[11]);
            }
            __syncthreads();
            pumping = sh[12];
        }
    }
    if (threadIdx.x == 0) {
        float max_comp = 0.0f, max_de = 0.0f;
        for (int i = 0; i < n_rows; ++i) {
            max_comp = fmaxf(max_comp, r_comp[i]);
            max_de = fmaxf(max_de, r_de[i]);
        }
        const float ceiling = args.ceilings[s];
        const float in_rms = acc.n_samples ? (float)sqrt(acc.sum_in / (double)acc.n_samples) : 0.0f;
        const float out_rms = acc.n_samples ? (float)sqrt(acc.sum_out / (double)acc.n_samples) : 0.0f;
        auto db = [](float v) { return __fmul_rn(20.0f, log10f(fmaxf(v, 1.0e-12f))); };
        AfChainMetrics m;
        m.input_sample_peak_db = db(acc.peak_in);
        m.input_rms_db = db(in_rms);
        m.output_sample_peak_db = db(acc.peak_out);
        m.pre_limiter_true_peak_db = db(acc.peak_pre_tp);
        m.output_true_peak_db = db(acc.peak_out_tp);
        m.output_rms_db = db(out_rms);
        m.limiter_effective_ceiling_db = ceiling;
        m.sample_headroom_db = __fsub_rn(ceiling, m.output_sample_peak_db);
        m.pre_limiter_true_peak_headroom_db = __fsub_rn(ceiling, m.pre_limiter_true_peak_db);
        m.true_peak_headroom_db = __fsub_rn(ceiling, m.output_true_peak_db);
        m.limiter_gain_reduction_db = acc.limiter_gr_db;
        m.true_peak_limiter_gain_reduction_db = acc.tp_gr_db;
        m.compressor_gain_reduction_db = max_comp;
        m.deesser_gain_reduction_db = max_de;
        m.compressor_gain_reduction_median_db = sh[2];
        m.compressor_gain_reduction_p95_db = sh[3];
        m.compressor_gain_reduction_active_ratio = sh[1];
        m.active_output_gain_db = sh[6];
        m.silence_output_gain_db = sh[8];
        m.silence_level_delta_db = sh[7];
        m.compressor_pumping_score_db = pumping;
        m.deesser_gain_reduction_median_db = sh[4];
        m.deesser_gain_reduction_p95_db = sh[5];
        m.analysis_block_ms = 20.0f;
        m.active_analysis_threshold_db = thr;
        m.non_finite_output = acc.non_finite;
        m.true_peak_limited_events = acc.events;
        m.active_analysis_block_count = (uint64_t)n_act;
        m.processed_samples = acc.n_samples;
        m.candidate_runtime_ms = 0.0;
        args.metrics[s] = m;
    }
}

// ---------------------------------------------------------------------------------------------------
// EQ magnitude response (dsp/eq.rs:514-527 over dsp/biquad.rs:184-205): one thread per (set, frequency)
// ---------------------------------------------------------------------------------------------------
__global__ void eq_response_kernel(const double* __restrict__ coeffs /* [set][40][5] */,
                                   const int* __restrict__ n_sections /* [set][10] band section counts */,
                                   const double* __restrict__ freqs, int n_freqs, int n_sets, double sample_rate,
                                   double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_freqs * n_sets) return;
    const int set = i / n_freqs;
    const double f = freqs[i % n_freqs];
    const double omega = 2.0 * 3.14159265358979323846264338327950288 * f / sample_rate;
    const double c1 = cos(omega), s1 = sin(omega), c2 = cos(2.0 * omega), s2 = sin(2.0 * omega);
    double total = 0.0;
    for (int band = 0; band < 10; ++band) {
        double band_sum = 0.0;
        const int ns = n_sections[set * 10 + band];
        for (int s = 0; s < ns; ++s) {
            const double* c = coeffs + ((size_t)set * kMaxSections + band * 4 + s) * 5;
            const double nr = c[0] + c[1] * c1 + c[2] * c2;
            const double ni = -c[1] * s1 - c[2] * s2;
            const double dr = 1.0 + c[3] * c1 + c[4] * c2;
            const double di = -c[3] * s1 - c[4] * s2;
            const double np = nr * nr + ni * ni;
            const double dp = dr * dr + di * di;
            const double mag = sqrt(np / fmax(dp, 1.0e-30));
            band_sum += 20.0 * log10(fmax(mag, 1.0e-10));
        }
        total += band_sum;
    }
    out[i] = total;
}

// Device-side synthetic passages for bench.py; kind 0 = speech-like, 1 = hot white noise.
__global__ void synth_kernel(float* out, size_t n_per, int n_passages, int kind, double sample_rate) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = n_per * (size_t)n_passages;
    if (i >= total) return;
    const int p = (int)(i / n_per);
    const size_t n = i % n_per;
    // counter-based LCG hash: state = seed + p, advanced n times is too slow; use a splitmix of (p, n)
    uint64_t z = 0x6a09e667f3bcc909ull + (uint64_t)p * 0x9e3779b97f4a7c15ull + (uint64_t)n * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z = z ^ (z >> 31);
    const float u = (float)((z >> 40) & 0xffffff) / 16777215.0f * 2.0f - 1.0f;
    float v;
    if (kind == 1) {
        v = u * 0.88f;  // ~5 % of samples above the -1.5 dBFS ceiling (0.841)
    } else {
        const float t = (float)((double)n / sample_rate);
        const float f0 = 110.0f + 13.75f * (float)(p % 9);
        const float env = 0.25f + 0.75f * fabsf(sinf(6.2831853f * 1.7f * t + 0.37f * (float)p));
        float h = 0.30f * sinf(6.2831853f * f0 * t) + 0.14f * sinf(6.2831853f * 2.0f * f0 * t) +
                  0.10f * sinf(6.2831853f * 3.0f * f0 * t) + 0.08f * sinf(6.2831853f * 2700.0f * t);
        const int gate = ((int)(t / 0.12f) % 5) == 2;
        v = 0.70f * (env * h + (gate ? 0.30f * sinf(6.2831853f * 7200.0f * t) : 0.0f) + u * 0.0126f);
    }
    out[i] = v;
}

// ---------------------------------------------------------------------------------------------------
// launch wrappers (called from afsim_api.cu)
// ---------------------------------------------------------------------------------------------------
size_t render_smem_bytes(int ring_rows) { return (size_t)(ring_rows + 96 + kTileMaxSlots) * kLanes * sizeof(float); }

cudaError_t launch_render(const RenderArgs& args, cudaStream_t stream) {
    const size_t smem = render_smem_bytes(args.ring_rows);
    cudaError_t err = cudaFuncSetAttribute(chain_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    chain_render_kernel<<<args.n_groups, 32, smem, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const FinalizeArgs& args, cudaStream_t stream) {
    const size_t smem = args.use_global_scratch ? 0 : (size_t)2 * args.n_pad * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    finalize_kernel<<<args.n_streams, 256, smem, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_eq_response(const double* coeffs, const int* n_sections, const double* freqs, int n_freqs,
                               int n_sets, double fs, double* out, cudaStream_t stream) {
    const int total = n_freqs * n_sets;
    eq_response_kernel<<<(total + 127) / 128, 128, 0, stream>>>(coeffs, n_sections, freqs, n_freqs, n_sets, fs, out);
    return cudaGetLastError();
}

cudaError_t launch_synth(float* out, size_t n_per, int n_passages, int kind, double fs, cudaStream_t stream) {
    const size_t total = n_per * (size_t)n_passages;
    synth_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(out, n_per, n_passages, kind, fs);
    return cudaGetLastError();
}

cudaError_t upload_eq_defaults(const double (*table)[5], cudaStream_t stream) {
    return cudaMemcpyToSymbolAsync(c_eq_default, table, sizeof(double) * 50, 0, cudaMemcpyHostToDevice, stream);
}

}  // namespace afsim
