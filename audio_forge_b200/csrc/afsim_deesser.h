// afsim_deesser.h -- the dynamic-EQ de-esser (rust-core/src/dsp/deesser.rs:405-547) cut into three serial
// recurrence kernels and two map kernels:
//     R_a  detector biquads (3 bands x high-pass + low-pass), band / broadband envelopes
//     M_b  levels in dB, voice level, narrowness, dominance, confidence targets   (4 log10, 3 sqrt, ~20 divisions)
//     R_c1 confidence / baseline / reduction smoothing, the target logic, the dynamic-EQ gain hysteresis
//     M_c2 the peaking-filter coefficients of the samples whose gain moved (exp10 + six divisions per band)
//     R_c3 the three time-varying peaking biquads
// The maps hold most of the arithmetic and run one thread per (stream, 2 samples); the serial kernels keep
// ~16 / ~13 / ~24 state values per stream in registers.  Per sample the operations and their order are those of
// DeEsser::process_sample, so the result is the same as a fused walk (tests/hostsim: bit-exact vs the oracle).
//
// Hand-off rings (f64, [row][stream]): R_a writes w0 = broadband envelope, w1..w3 = band envelopes; M_b
// overwrites w0 = voice dB, w1..w3 = band level dB and writes w4..w6 = clamped confidence targets.
#pragma once
#include "afsim_split.h"

namespace afsim {

constexpr int kDeMapGroup = 2;    // samples per thread of the de-esser map
constexpr int kDeRcDepth = 3;     // staging depth of R_c (8 staged streams: keep the shared-memory footprint small)
constexpr int kStateDeDetect = 16;  // state slots of R_a; R_c1's and R_c3's follow

// ---- R_a ------------------------------------------------------------------------------------------------------
struct DeEsserDetect {
    double dz[3][4];  // detector hp z1,z2, lp z1,z2
    double env[3];
    double broadband;

    AF_HD void init() {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dz[b][i] = 0.0;
            env[b] = 0.0;
        }
        broadband = 0.0;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
#pragma unroll
            for (int i = 0; i < 4; ++i) io.f64(dz[b][i]);
            io.f64(env[b]);
        }
        io.f64(broadband);
    }

    // x: chunk column (f32); w: ring columns at chunk start.  The first F samples of a render cross-fade the
    // detector biquads from their constructor coefficients (deesser.rs:64-73 -> dsp/biquad.rs:249-327).
    AF_HD void run(const float* x, double* w0, double* w1, double* w2, double* w3, size_t stride, int n0, int len,
                   int fade_total, const DeConst& k, const CandidateParams* p, Staging stg) {
        constexpr int U = kGroup;
        const double det_attack = k(DE_DET_ATTACK), det_release = k(DE_DET_RELEASE);
        Bq hp[3], lp[3];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            hp[b] = k.bq(DE_DET + 10 * b);
            lp[b] = k.bq(DE_DET + 10 * b + 5);
        }
        int t_head = 0;
        if (n0 < fade_total) {  // head of the render, sample by sample
            double pdz[3][4];
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int i = 0; i < 4; ++i) pdz[b][i] = 0.0;
            // (the crossfade always starts at n = 0 and chunks are at least F long: n0 == 0 here)
            for (; t_head < len && n0 + t_head < fade_total; ++t_head) {
                const int n = n0 + t_head;
                const float in = x[(size_t)t_head * stride];
                broadband = smooth_ar(broadband, (double)fabsf(in), det_attack, det_release);
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    const Bq hp0 = bq_from(p->de_det0[2 * b]);
                    const Bq lp0 = bq_from(p->de_det0[2 * b + 1]);
                    const float hp_out = (float)bq_step_fading((double)in, hp0, hp[b], dz[b][0], dz[b][1], pdz[b][0],
                                                               pdz[b][1], n, fade_total);
                    const float sc = (float)bq_step_fading((double)hp_out, lp0, lp[b], dz[b][2], dz[b][3], pdz[b][2],
                                                           pdz[b][3], n, fade_total);
                    if (n + 1 == fade_total) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) dz[b][i] = pdz[b][i];
                    }
                    env[b] = smooth_ar(env[b], (double)fabsf(sc), det_attack, det_release);
                }
                w0[(size_t)t_head * stride] = broadband;
                w1[(size_t)t_head * stride] = env[0];
                w2[(size_t)t_head * stride] = env[1];
                w3[(size_t)t_head * stride] = env[2];
            }
        }
        // steady state: staged tiles
        const float* xs = x + (size_t)t_head * stride;
        double* o0 = w0 + (size_t)t_head * stride;
        double* o1 = w1 + (size_t)t_head * stride;
        double* o2 = w2 + (size_t)t_head * stride;
        double* o3 = w3 + (size_t)t_head * stride;
        const int m = len - t_head;
        const StageRing<float> sx = stg.ring<float>();
        auto issue = [&](int kt, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = kt * U;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (FULL || t0 + u < m) sx.fetch(kt, u, xs + (size_t)(t0 + u) * stride);
        };
        auto body = [&](int kt, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = kt * U;
            const int valid = FULL ? U : m - t0;
            float xin[U];
#pragma unroll
            for (int u = 0; u < U; ++u) xin[u] = (FULL || u < valid) ? sx.get(kt, u, xs + (size_t)(t0 + u) * stride) : 0.0f;
            double ob[U], o_e0[U], o_e1[U], o_e2[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || u < valid) {
                    const float in = xin[u];
                    broadband = smooth_ar(broadband, (double)fabsf(in), det_attack, det_release);
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const float hp_out = (float)bq_step((double)in, hp[b], dz[b][0], dz[b][1]);
                        const float sc = (float)bq_step((double)hp_out, lp[b], dz[b][2], dz[b][3]);
                        env[b] = smooth_ar(env[b], (double)fabsf(sc), det_attack, det_release);
                    }
                }
                ob[u] = broadband;
                o_e0[u] = env[0];
                o_e1[u] = env[1];
                o_e2[u] = env[2];
            }
            store_tile(o0 + (size_t)t0 * stride, stride, valid, ob);
            store_tile(o1 + (size_t)t0 * stride, stride, valid, o_e0);
            store_tile(o2 + (size_t)t0 * stride, stride, valid, o_e1);
            store_tile(o3 + (size_t)t0 * stride, stride, valid, o_e2);
        };
        pipelined_tiles(m, issue, body);
    }
};

// ---- M_b ------------------------------------------------------------------------------------------------------
AF_HD void deesser_levels(double* w0, double* w1, double* w2, double* w3, double* w4, double* w5, double* w6, size_t stride,
                          int valid) {
    constexpr int G = kDeMapGroup;
    double bb[G], e0[G], e1[G], e2[G];
    load_tile((const double*)w0, stride, valid, bb);
    load_tile((const double*)w1, stride, valid, e0);
    load_tile((const double*)w2, stride, valid, e1);
    load_tile((const double*)w3, stride, valid, e2);
    double voice[G], lv0[G], lv1[G], lv2[G], c0[G], c1[G], c2[G];
#pragma unroll
    for (int u = 0; u < G; ++u) {
        const double env[3] = {e0[u], e1[u], e2[u]};
        double level_db[3];
        double total_env = 0.0, max_env = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            total_env += env[b];
            max_env = fmax(max_env, env[b]);
            level_db[b] = lin_to_db(env[b], 1e-10);
        }
        const double voice_level = fmax(bb[u] - total_env * 0.6, 1e-8);
        const double voice_db = lin_to_db(voice_level, 1e-10);
        const double narrowness = total_env > 1e-10 ? max_env / total_env : 0.0;
        double conf[3];
        const DeConfShared shared = de_confidence_shared(voice_db, narrowness);
        const AfDivisor by_max = af_divisor(max_env);  // one refined reciprocal for the three dominance quotients
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double dominance = max_env > 1e-10 ? sqrt(af_div(env[b], by_max)) : 0.0;
            conf[b] = clampd(de_confidence_target(level_db[b], voice_db, shared) * dominance, 0.0, 1.0);
        }
        voice[u] = voice_db;
        lv0[u] = level_db[0];
        lv1[u] = level_db[1];
        lv2[u] = level_db[2];
        c0[u] = conf[0];
        c1[u] = conf[1];
        c2[u] = conf[2];
    }
    store_tile(w0, stride, valid, voice);
    store_tile(w1, stride, valid, lv0);
    store_tile(w2, stride, valid, lv1);
    store_tile(w3, stride, valid, lv2);
    store_tile(w4, stride, valid, c0);
    store_tile(w5, stride, valid, c1);
    store_tile(w6, stride, valid, c2);
}

// ---- R_c1 / M_c2 / R_c3 -------------------------------------------------------------------------------------------
// The second half of DeEsser::process_sample (deesser.rs:452-547) in three kernels:
//   R_c1  serial, light: confidence / baseline smoothing, the reduction targets and their cap, reduction smoothing
//         and the 0.001 dB gain hysteresis of set_gain_db_immediate.  Per sample it hands on a 3-bit mask of the
//         bands whose peaking filter is rebuilt (w0) and the gain they are rebuilt with (w4..w6).
//   M_c2  map: the rebuilt coefficients (exp10, six divisions per band) -- the expensive part, a pure function of
//         the gain -- for the flagged (sample, band) pairs: b0 b1 b2 a2 (a1 == b1) into w1..w3 / w4..w6 / w7..w12.
//   R_c3  serial, light: the three time-varying biquads (with the configuration crossfade of the first F samples
//         and its cancellation by the first rebuild), in place on the signal.
// Cutting the serial chain in two and moving the rebuilds to a map takes the de-esser off the wavefront's critical
// path: R_c was ~515 instructions per sample on one warp per 32 streams.

// The constants R_c1 reads every sample, loaded once per chunk.
struct DeApplyConst {
    double det_attack, det_release, max_red, attack, release;
    double base_fall, base_rise, base_inactive, conf_floor, trigger, slope, cap;      // auto mode
    double threshold, ratio_thr, ratio_factor, manual_cap;                            // manual mode
    AF_HD void load(const DeConst& k) {
        det_attack = k(DE_DET_ATTACK);
        det_release = k(DE_DET_RELEASE);
        max_red = k(DE_MAX_RED);
        attack = k(DE_ATTACK);
        release = k(DE_RELEASE);
        base_fall = k(DE_BASE_FALL);
        base_rise = k(DE_BASE_RISE);
        base_inactive = k(DE_BASE_INACTIVE);
        conf_floor = k(DE_CONF_FLOOR);
        trigger = k(DE_TRIGGER);
        slope = k(DE_SLOPE);
        cap = k(DE_CAP);
        threshold = k(DE_THRESHOLD);
        ratio_thr = k(DE_RATIO_THR);
        ratio_factor = k(DE_RATIO_FACTOR);
        manual_cap = k(DE_MANUAL_CAP);
    }
};

constexpr int kStateDeTargets = 16;  // state slots of R_c1 (after R_a's); R_c3's follow

struct DeEsserTargets {  // R_c1
    double conf[3], base[3], red[3], built_gain[3];
    double current;
    bool auto_mode;

    AF_HD void init(const CandidateParams& p) {
#pragma unroll
        for (int b = 0; b < 3; ++b) conf[b] = base[b] = red[b] = built_gain[b] = 0.0;
        current = 0.0;
        auto_mode = (p.flags & LF_DE_AUTO) != 0;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            io.f64(conf[b]);
            io.f64(base[b]);
            io.f64(red[b]);
            io.f64(built_gain[b]);
        }
        io.f64(current);
    }

    // One sample: returns the rebuild mask, gains[b] = the new gain of a flagged band.
    AF_HD unsigned sample(double voice_db, const double (&level_db)[3], const double (&conf_target)[3], const DeApplyConst& k,
                          double conf_lo, const AfDivisor& conf_div, double (&gains)[3]) {
        double target[3];
        double target_sum = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double ratio_db = fmax(level_db[b] - voice_db, 0.0);
            conf[b] = smooth_ar(conf[b], conf_target[b], k.det_attack, k.det_release);
            double tr = 0.0;
            if (auto_mode) {
                const bool voice_active = voice_db > -55.0 || level_db[b] > -55.0;
                if (voice_active) {
                    const double base_target = clampd(ratio_db * 0.45, 0.0, 24.0);
                    const double c = base_target < base[b] ? k.base_fall : k.base_rise;
                    base[b] = c * base[b] + (1.0 - c) * base_target;
                } else {
                    base[b] *= k.base_inactive;
                }
                const double conf_gain = clampd(af_div(conf[b] - conf_lo, conf_div), 0.0, 1.0);  // norm_range(conf, floor, 1)
                const double over_db = fmax(ratio_db - base[b] - k.trigger, 0.0);
                tr = clampd(over_db * k.slope * conf_gain, 0.0, k.cap);
            } else if (level_db[b] > k.threshold) {
                const double level_over = level_db[b] - k.threshold;
                const double ratio_over = ratio_db - k.ratio_thr;
                if (ratio_over > 0.0) {
                    const double over_db = fmin(level_over, ratio_over);
                    const double conf_gain = clampd(af_div(conf[b] - conf_lo, conf_div), 0.0, 1.0);
                    tr = clampd(k.ratio_factor * over_db * conf_gain, 0.0, k.manual_cap);
                }
            }
            target[b] = tr;
            target_sum += tr;
        }
        if (target_sum > k.max_red && target_sum > 0.0) {
            const double scale = k.max_red / target_sum;
#pragma unroll
            for (int b = 0; b < 3; ++b) target[b] *= scale;
        }
        unsigned mask = 0;
        double total_red = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            red[b] = smooth_ar(red[b], target[b], k.attack, k.release);
            total_red += red[b];
            const double dyn_gain = -red[b];
            gains[b] = dyn_gain;
            if (fabs(built_gain[b] - dyn_gain) > 0.001) {  // set_gain_db_immediate rebuilds the filter
                built_gain[b] = dyn_gain;
                mask |= 1u << b;
            }
        }
        current = fmin(total_red, k.max_red);
        return mask;
    }

    // in[0..6]: columns at chunk start (pitch in_stride) of voice dB, 3 level dB, 3 confidence targets -- the stream's
    // own rings, or the shared front of its (passage, detector) pair; on return w[0] holds the rebuild masks and
    // w[4..6] the gains (the stream's rings, pitch stride)
    AF_HD void run(const double* const (&in)[7], size_t in_stride, double* const (&w)[7], size_t stride, int n0, int len,
                   const DeConst& table, BlockClock clk, float* rows_de, Staging stg) {
        constexpr int U = kGroup;
        DeApplyConst k;
        k.load(table);
        const double conf_lo = auto_mode ? k.conf_floor : 0.22;
        const AfDivisor conf_div = af_divisor(1.0 - conf_lo);
        StageRing<double, kDeRcDepth> sw[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) sw[i] = stg.ring<double, kDeRcDepth>();
        auto issue = [&](int kt, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = kt * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || t0 + u < len) {
                    const size_t o = (size_t)(t0 + u) * in_stride;
#pragma unroll
                    for (int i = 0; i < 7; ++i) sw[i].fetch(kt, u, in[i] + o);
                }
            }
        };
        // (a phased tile -- recurrences / feed-forward targets / recurrences -- was measured slower: it spills)
        auto body = [&](int kt, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = kt * U;
            const int valid = FULL ? U : len - t0;
#pragma unroll 2
            for (int u = 0; u < U; ++u) {
                if (FULL || u < valid) {
                    const size_t o = (size_t)(t0 + u) * stride, oi = (size_t)(t0 + u) * in_stride;
                    const double voice_db = sw[0].get(kt, u, in[0] + oi);
                    const double lv[3] = {sw[1].get(kt, u, in[1] + oi), sw[2].get(kt, u, in[2] + oi), sw[3].get(kt, u, in[3] + oi)};
                    const double ct[3] = {sw[4].get(kt, u, in[4] + oi), sw[5].get(kt, u, in[5] + oi), sw[6].get(kt, u, in[6] + oi)};
                    double gains[3];
                    const unsigned mask = sample(voice_db, lv, ct, k, conf_lo, conf_div, gains);
                    w[0][o] = (double)mask;
                    if (mask & 1u) w[4][o] = gains[0];  // M_c2 reads the gain of flagged (sample, band) pairs only
                    if (mask & 2u) w[5][o] = gains[1];
                    if (mask & 4u) w[6][o] = gains[2];
                    if (clk.at_end(n0 + t0 + u)) {  // block-end meter sample (block_processor.rs:129-133)
                        rows_de[(size_t)clk.blk * stride] = (float)current;
                        clk.advance();
                    }
                }
            }
        };
        pipelined_tiles_depth<kDeRcDepth>(len, issue, body);
    }
};
constexpr size_t kDeRc1StagingBytesPerLane = (size_t)kDeRcDepth * 8 * (7 * 8);

// ---- R_c1 cut three ways (few-stream batches) ---------------------------------------------------------------------------------
// In a few-stream batch R_c1 is the wavefront's slowest stage: one warp per 32 streams walks ~440 dependent-issue
// instructions per sample, of which only the three smoothers (confidence, baseline, reduction) and the rebuild
// hysteresis carry state from sample to sample.  The cut keeps those in two short serial kernels and moves the
// feed-forward middle (normalised confidence, over-threshold, target, the max-reduction rescale: two divisions per
// sample) into a map over (stream, sample group):
//   R_c1a  conf[b], base[b] smoothed      (voice, level[3], conf target[3]) -> w7..w9 = conf, w10..w12 = base
//   M_c1b  target[b], rescaled            (voice, level[3], w7..w12)        -> w7..w9 = target
//   R_c1c  red[b] smoothed, hysteresis    (w7..w9)                          -> w0 = rebuild mask, w4..w6 = gains
// w7..w12 are free at this point (M_c2 writes its coefficients there afterwards).  Every expression is the one
// `DeEsserTargets::sample` evaluates, in the same order: the results are identical bit for bit.  The state table
// keeps R_c1's layout (conf, base, red, built per band, then current); each kernel parks its own slots.
struct DeEsserConfBase {  // R_c1a
    double conf[3], base[3];
    bool auto_mode;
    AF_HD void init(const CandidateParams& p) {
#pragma unroll
        for (int b = 0; b < 3; ++b) conf[b] = base[b] = 0.0;
        auto_mode = (p.flags & LF_DE_AUTO) != 0;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            io.f64(conf[b]);
            io.f64(base[b]);
            io.skip(2);
        }
    }
    // in[0..6] as for DeEsserTargets::run; out[0..5]: the stream's w7..w12 columns at chunk start (pitch stride)
    AF_HD void run(const double* const (&in)[7], size_t in_stride, double* const (&out)[6], size_t stride, int len,
                   const DeConst& table, Staging stg) {
        constexpr int U = kGroup;
        DeApplyConst k;
        k.load(table);
        StageRing<double, kDeRcDepth> sw[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) sw[i] = stg.ring<double, kDeRcDepth>();
        auto issue = [&](int kt, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = kt * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || t0 + u < len) {
                    const size_t o = (size_t)(t0 + u) * in_stride;
#pragma unroll
                    for (int i = 0; i < 7; ++i) sw[i].fetch(kt, u, in[i] + o);
                }
            }
        };
        auto body = [&](int kt, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = kt * U;
            const int valid = FULL ? U : len - t0;
#pragma unroll 2
            for (int u = 0; u < U; ++u) {
                if (FULL || u < valid) {
                    const size_t o = (size_t)(t0 + u) * stride, oi = (size_t)(t0 + u) * in_stride;
                    const double voice_db = sw[0].get(kt, u, in[0] + oi);
                    const double lv[3] = {sw[1].get(kt, u, in[1] + oi), sw[2].get(kt, u, in[2] + oi), sw[3].get(kt, u, in[3] + oi)};
                    const double ct[3] = {sw[4].get(kt, u, in[4] + oi), sw[5].get(kt, u, in[5] + oi), sw[6].get(kt, u, in[6] + oi)};
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        conf[b] = smooth_ar(conf[b], ct[b], k.det_attack, k.det_release);
                        if (auto_mode) {
                            const double ratio_db = fmax(lv[b] - voice_db, 0.0);
                            const bool voice_active = voice_db > -55.0 || lv[b] > -55.0;
                            if (voice_active) {
                                const double base_target = clampd(ratio_db * 0.45, 0.0, 24.0);
                                const double c = base_target < base[b] ? k.base_fall : k.base_rise;
                                base[b] = c * base[b] + (1.0 - c) * base_target;
                            } else {
                                base[b] *= k.base_inactive;
                            }
                        }
                        out[b][o] = conf[b];
                        out[3 + b][o] = base[b];
                    }
                }
            }
        };
        pipelined_tiles_depth<kDeRcDepth>(len, issue, body);
    }
};

constexpr size_t kDeRc1cStagingBytesPerLane = (size_t)kDeRcDepth * 8 * (3 * 8);

// M_c1b: in[0..3] voice / level columns at the group's first sample (pitch in_stride), w[0..5] the stream's w7..w12 there
constexpr int kDeTargetGroup = 2;
AF_HD void deesser_targets(const double* const (&in)[4], size_t in_stride, double* const (&w)[6], size_t stride, int valid,
                           const DeConst& table, bool auto_mode) {
    constexpr int G = kDeTargetGroup;
    DeApplyConst k;
    k.load(table);
    const double conf_lo = auto_mode ? k.conf_floor : 0.22;
    const AfDivisor conf_div = af_divisor(1.0 - conf_lo);
    double voice[G], lv[3][G], conf[3][G], base[3][G], target[3][G];
    load_tile(in[0], in_stride, valid, voice);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        load_tile(in[1 + b], in_stride, valid, lv[b]);
        load_tile((const double*)w[b], stride, valid, conf[b]);
        load_tile((const double*)w[3 + b], stride, valid, base[b]);
    }
#pragma unroll
    for (int u = 0; u < G; ++u) {
        double target_sum = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double ratio_db = fmax(lv[b][u] - voice[u], 0.0);
            double tr = 0.0;
            if (auto_mode) {
                const double conf_gain = clampd(af_div(conf[b][u] - conf_lo, conf_div), 0.0, 1.0);
                const double over_db = fmax(ratio_db - base[b][u] - k.trigger, 0.0);
                tr = clampd(over_db * k.slope * conf_gain, 0.0, k.cap);
            } else if (lv[b][u] > k.threshold) {
                const double level_over = lv[b][u] - k.threshold;
                const double ratio_over = ratio_db - k.ratio_thr;
                if (ratio_over > 0.0) {
                    const double over_db = fmin(level_over, ratio_over);
                    const double conf_gain = clampd(af_div(conf[b][u] - conf_lo, conf_div), 0.0, 1.0);
                    tr = clampd(k.ratio_factor * over_db * conf_gain, 0.0, k.manual_cap);
                }
            }
            target[b][u] = tr;
            target_sum += tr;
        }
        if (target_sum > k.max_red && target_sum > 0.0) {
            const double scale = k.max_red / target_sum;
#pragma unroll
            for (int b = 0; b < 3; ++b) target[b][u] *= scale;
        }
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) store_tile(w[b], stride, valid, target[b]);
}

struct DeEsserReduction {  // R_c1c
    double red[3], built_gain[3];
    double current;
    AF_HD void init() {
#pragma unroll
        for (int b = 0; b < 3; ++b) red[b] = built_gain[b] = 0.0;
        current = 0.0;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            io.skip(2);
            io.f64(red[b]);
            io.f64(built_gain[b]);
        }
        io.f64(current);
    }
    // tg[0..2]: the stream's w7..w9 columns at chunk start; mask_out = w0, gain_out[0..2] = w4..w6 (pitch stride)
    AF_HD void run(const double* const (&tg)[3], double* mask_out, double* const (&gain_out)[3], size_t stride, int n0, int len,
                   const DeConst& table, BlockClock clk, float* rows_de, Staging stg) {
        constexpr int U = kGroup;
        const double attack = table(DE_ATTACK), release = table(DE_RELEASE), max_red = table(DE_MAX_RED);
        StageRing<double, kDeRcDepth> sw[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) sw[i] = stg.ring<double, kDeRcDepth>();
        auto issue = [&](int kt, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = kt * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || t0 + u < len) {
                    const size_t o = (size_t)(t0 + u) * stride;
#pragma unroll
                    for (int i = 0; i < 3; ++i) sw[i].fetch(kt, u, tg[i] + o);
                }
            }
        };
        auto body = [&](int kt, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = kt * U;
            const int valid = FULL ? U : len - t0;
#pragma unroll 2
            for (int u = 0; u < U; ++u) {
                if (FULL || u < valid) {
                    const size_t o = (size_t)(t0 + u) * stride;
                    const double target[3] = {sw[0].get(kt, u, tg[0] + o), sw[1].get(kt, u, tg[1] + o), sw[2].get(kt, u, tg[2] + o)};
                    unsigned mask = 0;
                    double total_red = 0.0;
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        red[b] = smooth_ar(red[b], target[b], attack, release);
                        total_red += red[b];
                        const double dyn_gain = -red[b];
                        if (fabs(built_gain[b] - dyn_gain) > 0.001) {  // set_gain_db_immediate rebuilds the filter
                            built_gain[b] = dyn_gain;
                            mask |= 1u << b;
                            gain_out[b][o] = dyn_gain;  // M_c2 reads the gain of flagged (sample, band) pairs only
                        }
                    }
                    current = fmin(total_red, max_red);
                    mask_out[o] = (double)mask;
                    if (clk.at_end(n0 + t0 + u)) {  // block-end meter sample (block_processor.rs:129-133)
                        rows_de[(size_t)clk.blk * stride] = (float)current;
                        clk.advance();
                    }
                }
            }
        };
        pipelined_tiles_depth<kDeRcDepth>(len, issue, body);
    }
};

// ---- M_c2: coefficients of the rebuilt filters ------------------------------------------------------------------------
// coefficient rings of band b: w[kDeCoefRing[b][i]], i = b0 b1 b2 a2
AF_HD int de_coef_ring(int band, int i) { return band == 0 ? (i < 3 ? 1 + i : 7) : (band == 1 ? 8 + i : (i < 1 ? 12 : 3 + i)); }
// band 0: w1 w2 w3 w7 ; band 1: w8 w9 w10 w11 ; band 2: w12 w4 w5 w6 (the gains in w4..w6 are read before they are overwritten)
constexpr int kDeRebuildGroup = 2;
AF_HD void deesser_rebuild(double* const (&w)[13], size_t stride, int valid, const DeConst& k) {
    constexpr int G = kDeRebuildGroup;
    double mask[G];
    load_tile((const double*)w[0], stride, valid, mask);
#pragma unroll
    for (int u = 0; u < G; ++u) {
        if (u >= valid) continue;
        const unsigned m = (unsigned)mask[u];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            if (m & (1u << b)) {  // the gain rings hold a value for the flagged pairs only (R_c1c), and only those are fetched
                const double gain = w[4 + b][(size_t)u * stride];
                const Bq c = design_peaking(k(DE_DYN_COS + b), k(DE_DYN_ALPHA + b), gain);
                w[de_coef_ring(b, 0)][(size_t)u * stride] = c.b0;
                w[de_coef_ring(b, 1)][(size_t)u * stride] = c.b1;
                w[de_coef_ring(b, 2)][(size_t)u * stride] = c.b2;
                w[de_coef_ring(b, 3)][(size_t)u * stride] = c.a2;
            }
        }
    }
}

// ---- R_c3: the time-varying dynamic EQ ------------------------------------------------------------------------------------
constexpr int kDeRc3Depth = 3;
constexpr int kDeRc3Look = kDeRc3Depth;  // tiles the rebuild masks are staged ahead of the coefficients they gate
struct DeEsserFilter {
    Bq dyn[3];        // live dynamic-EQ coefficients
    double yz[3][2];  // dynamic-EQ state
    bool cancel[3];   // set_gain_db_immediate cancelled the configuration crossfade

    AF_HD void init(const CandidateParams& p) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            dyn[b] = bq_from(p.de_dyn0[b]);
            yz[b][0] = yz[b][1] = 0.0;
            cancel[b] = false;
        }
    }
    template <class IO>
    AF_HD void sync(IO& io) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            io.f64(dyn[b].b0);
            io.f64(dyn[b].b1);
            io.f64(dyn[b].b2);
            io.f64(dyn[b].a1);
            io.f64(dyn[b].a2);
            io.f64(yz[b][0]);
            io.f64(yz[b][1]);
            io.flag(cancel[b]);
        }
    }

    // xin: input signal column at chunk start (pitch xin_stride): the stream's own buf_a column (in place) or the
    // shared input stage output of its passage; x: output column (pitch stride)
    AF_HD void run(const float* xin, size_t xin_stride, float* x, double* const (&w)[13], size_t stride, int n0, int len,
                   int fade_total, const CandidateParams* p, Staging stg) {
        constexpr int U = kGroup;
        int t_head = 0;
        if (n0 < fade_total) {  // head of the render: the configuration crossfade (deesser.rs:312-327), sample by sample
            double pyz[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
            for (; t_head < len && n0 + t_head < fade_total; ++t_head) {
                const int n = n0 + t_head;
                const size_t o = (size_t)t_head * stride;
                const unsigned mask = (unsigned)w[0][o];
                float processed = xin[(size_t)t_head * xin_stride];
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    if (mask & (1u << b)) {
                        dyn[b].b0 = w[de_coef_ring(b, 0)][o];
                        dyn[b].b1 = dyn[b].a1 = w[de_coef_ring(b, 1)][o];
                        dyn[b].b2 = w[de_coef_ring(b, 2)][o];
                        dyn[b].a2 = w[de_coef_ring(b, 3)][o];
                        cancel[b] = true;
                    }
                    double y;
                    if (!cancel[b]) {
                        const Bq pend = bq_from(p->de_dyn1[b]);
                        y = bq_step_fading((double)processed, dyn[b], pend, yz[b][0], yz[b][1], pyz[b][0], pyz[b][1], n, fade_total);
                        if (n + 1 == fade_total) {
                            dyn[b] = pend;
                            yz[b][0] = pyz[b][0];
                            yz[b][1] = pyz[b][1];
                        }
                    } else {
                        y = bq_step((double)processed, dyn[b], yz[b][0], yz[b][1]);
                    }
                    processed = (float)y;
                }
                x[o] = processed;
            }
        }
        float* xs = x + (size_t)t_head * stride;
        const float* xis = xin + (size_t)t_head * xin_stride;
        const double* ws[13];
#pragma unroll
        for (int i = 0; i < 13; ++i) ws[i] = w[i] + (size_t)t_head * stride;
        const int m = len - t_head;
        // Only the flagged (sample, band) pairs carry coefficients (M_c2 wrote nothing else), so only those are fetched: the
        // rebuild masks are staged kDeRc3Look tiles AHEAD of the coefficients they gate (their copies ride in the commit group
        // of tile kt - Look, which has landed when tile kt is issued: pipelined_tiles_depth waits for group k - 1 before it
        // issues tile k + Depth - 1, and Look = Depth); the first Look tiles of a chunk read their masks through.  Fetching all
        // 13 rings for every sample made this serial kernel HBM bound (1.28 GB per launch at 5.9 TB/s for 8192 streams).
        constexpr int kLook = kDeRc3Look;
        const StageRing<float, kDeRc3Depth> sx = stg.ring<float, kDeRc3Depth>();
        const StageRing<double, kDeRc3Depth + kLook> smask = stg.ring<double, kDeRc3Depth + kLook>();
        StageRing<double, kDeRc3Depth> sw[13];  // sw[0] unused: the mask has its own, deeper ring
#pragma unroll
        for (int i = 1; i < 13; ++i) sw[i] = stg.ring<double, kDeRc3Depth>();
        sw[0] = sw[1];
        auto issue = [&](int kt, auto full) {
            constexpr bool FULL = decltype(full)::value;
            if (!sx.on) return;  // direct mode: nothing is staged
            const int t0 = kt * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || t0 + u < m) {
                    const size_t o = (size_t)(t0 + u) * stride;
                    sx.fetch(kt, u, xis + (size_t)(t0 + u) * xin_stride);
                    unsigned mask;
                    if (kt < kLook) {
                        mask = (unsigned)ws[0][o];
                        smask.fetch(kt, u, ws[0] + o);
                    } else {
                        mask = (unsigned)*smask.at(kt, u);
                    }
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        if (mask & (1u << b)) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) sw[de_coef_ring(b, i)].fetch(kt, u, ws[de_coef_ring(b, i)] + o);
                        }
                    }
                    const int ta = t0 + kLook * U + u;
                    if (ta < m) smask.fetch(kt + kLook, u, ws[0] + (size_t)ta * stride);
                }
            }
        };
        auto body = [&](int kt, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = kt * U;
            const int valid = FULL ? U : m - t0;
            float y[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                y[u] = 0.0f;
                if (FULL || u < valid) {
                    const size_t o = (size_t)(t0 + u) * stride;
                    const unsigned mask = (unsigned)smask.get(kt, u, ws[0] + o);
                    float processed = sx.get(kt, u, xis + (size_t)(t0 + u) * xin_stride);
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        if (mask & (1u << b)) {
                            dyn[b].b0 = sw[de_coef_ring(b, 0)].get(kt, u, ws[de_coef_ring(b, 0)] + o);
                            dyn[b].b1 = dyn[b].a1 = sw[de_coef_ring(b, 1)].get(kt, u, ws[de_coef_ring(b, 1)] + o);
                            dyn[b].b2 = sw[de_coef_ring(b, 2)].get(kt, u, ws[de_coef_ring(b, 2)] + o);
                            dyn[b].a2 = sw[de_coef_ring(b, 3)].get(kt, u, ws[de_coef_ring(b, 3)] + o);
                            cancel[b] = true;
                        }
                        processed = (float)bq_step((double)processed, dyn[b], yz[b][0], yz[b][1]);
                    }
                    y[u] = processed;
                }
            }
            store_tile(xs + (size_t)t0 * stride, stride, valid, y);
        };
        pipelined_tiles_depth<kDeRc3Depth>(m, issue, body);
    }
};
constexpr size_t kDeRc3StagingBytesPerLane = (size_t)kDeRc3Depth * 8 * (4 + 12 * 8) + (size_t)(kDeRc3Depth + kDeRc3Look) * 8 * 8;

}  // namespace afsim
