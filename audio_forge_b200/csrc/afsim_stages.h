// afsim_stages.h -- the per-stream stage arithmetic of the chain render kernels.
//
// Every stage is a struct that keeps ONE stream's recurrence state (a CUDA thread holds it in
// registers while it walks a chunk of samples) with
//     init(...)          the reference's freshly constructed + configured state,
//     sync(io)           park / restore the state in the stream-minor state table between chunks,
//     run(col, ...)      advance over one chunk held in a time-major work buffer column.
// Arithmetic follows the reference's evaluation order exactly: f64 state and coefficients, f32
// hand-off between stages, no multiply-add contraction (compile with -fmad=false /
// -ffp-contract=off) except the explicitly fused true-peak FIR.
//
// The code is host/device portable so that tests/hostsim can execute the very same stage bodies on
// the CPU (a test harness that checks chunking and state parking without a GPU; the product only
// ever runs them inside the CUDA kernels of afsim_kernels.cu).
//
// Reference citations are relative to rust-core/src/.
#pragma once
#include <math.h>
#include <stdint.h>

#include "afsim_params.h"

#if defined(__CUDACC__)
#define AF_HD __host__ __device__ __forceinline__
#else
#define AF_HD inline
#endif

#include "afsim_math.h"

namespace afsim {

// ---- numeric helpers -----------------------------------------------------------------------------------
AF_HD float af_fmaf(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
AF_HD double af_exp10(double x) { return af_exp10_fast(x); }  // afsim_math.h: the library's algorithm, constant-bank operands
AF_HD float af_log10_f32(float v) {  // f32::log10
#if defined(__CUDA_ARCH__)
    return (float)log10((double)v);
#else
    return log10f(v);
#endif
}
AF_HD bool af_finite(float v) { return fabsf(v) <= 3.402823466e+38f; }  // false for inf and NaN
AF_HD double lin_to_db(double v, double floor_) { return 20.0 * af_log10(fmax(fabs(v), floor_)); }  // dsp/util.rs:17-20
AF_HD double db_to_lin(double db) { return af_exp10(af_div_const(db, 20.0, 0.05)); }                 // dsp/util.rs:11-14
AF_HD float lin_to_db_f32(float v) { return 20.0f * af_log10_f32(fmaxf(v, 1.0e-12f)); }            // python_api.rs:54-56
AF_HD double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
AF_HD float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
AF_HD double smooth_ar(double prev, double in, double attack, double release) {  // dsp/deesser.rs:147-154
    const double c = in > prev ? attack : release;
    return c * prev + (1.0 - c) * in;
}
AF_HD double norm_range(double v, double s, double e) { return clampd((v - s) / (e - s), 0.0, 1.0); }
AF_HD double lerpd(double a, double b, double t) { return a + (b - a) * t; }

// One stream's column of a time-major work buffer: element t of the current chunk.
struct Col {
    float* base;
    size_t stride;
    AF_HD float get(int t) const { return base[(size_t)t * stride]; }
    AF_HD void set(int t, float v) const { base[(size_t)t * stride] = v; }
};

// Parks (STORE) or restores stage state: slot k of stream s lives at table[k * stride + s].
template <bool STORE>
struct StateIO {
    double* p;
    size_t stride;
    AF_HD void f64(double& v) {
        if (STORE) *p = v; else v = *p;
        p += stride;
    }
    AF_HD void f32(float& v) {
        if (STORE) *p = (double)v; else v = (float)*p;
        p += stride;
    }
    AF_HD void u32(uint32_t& v) {
        if (STORE) *p = (double)v; else v = (uint32_t)*p;
        p += stride;
    }
    AF_HD void flag(bool& v) {
        if (STORE) *p = v ? 1.0 : 0.0; else v = *p != 0.0;
        p += stride;
    }
    AF_HD void skip(int n) { p += (size_t)n * stride; }  // slots another kernel of the same stage owns
};

// Analysis-block clock (python_api.rs:512-513): block b ends after sample (b+1)*block-1 or T-1.
struct BlockClock {
    int blk, next_end, block, total;
    AF_HD void init(int block_samples, int n_samples, int n0) {
        block = block_samples;
        total = n_samples;
        blk = n0 / block_samples;
        next_end = (blk + 1) * block_samples;
    }
    AF_HD bool at_end(int n) const { return n + 1 == next_end || n + 1 == total; }
    AF_HD int block_len(int n) const { return n + 1 - (next_end - block); }
    // true when neither the current analysis block nor the signal ends within samples [n, n + count)
    AF_HD bool ends_after(int n, int count) const { return next_end - n > count && total - n > count; }
    AF_HD void advance() {
        blk += 1;
        next_end += block;
    }
};

struct Bq {  // one DF2T section, dsp/biquad.rs:262-274
    double b0, b1, b2, a1, a2;
};
AF_HD Bq bq_from(const double* c) {
    Bq r;
    r.b0 = c[0];
    r.b1 = c[1];
    r.b2 = c[2];
    r.a1 = c[3];
    r.a2 = c[4];
    return r;
}
AF_HD Bq bq_from_strided(const double* c, size_t stride) {
    Bq r;
    r.b0 = c[0];
    r.b1 = c[stride];
    r.b2 = c[2 * stride];
    r.a1 = c[3 * stride];
    r.a2 = c[4 * stride];
    return r;
}
AF_HD double bq_step(double x, const Bq& c, double& z1, double& z2) {
    const double y = c.b0 * x + z1;
    z1 = c.b1 * x - c.a1 * y + z2;
    z2 = c.b2 * x - c.a2 * y;
    return y;
}
// A section inside its coefficient crossfade (dsp/biquad.rs:290-327): the `active` filter fades into
// `pending` over fade_total samples; n = samples already faded.
AF_HD double bq_step_fading(double x, const Bq& active, const Bq& pending, double& z1, double& z2, double& pz1,
                            double& pz2, int n, int fade_total) {
    const double ya = bq_step(x, active, z1, z2);
    const double yp = bq_step(x, pending, pz1, pz2);
    const double fade = (double)(n + 1) / (double)fade_total;
    return ya * (1.0 - fade) + yp * fade;
}

// ---- input: sanitize, optional DC block (f32) + 80 Hz high-pass (f64), input statistics -------------------
// python_api.rs:517-524,549-551 ; audio/processor/routing.rs:826-843
struct InputStage {
    double sum_in, blk_in;
    double z1, z2;
    float peak_in;
    float x1, y1;

    AF_HD void init() {
        sum_in = blk_in = 0.0;
        z1 = z2 = 0.0;
        peak_in = 0.0f;
        x1 = y1 = 0.0f;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
        io.f64(sum_in);
        io.f64(blk_in);
        io.f64(z1);
        io.f64(z2);
        io.f32(peak_in);
        io.f32(x1);
        io.f32(y1);
    }
    // rows_in: row.0 table of this stream, element b at rows_in[b * stride]
    template <bool DC_HP>
    AF_HD void step(float v, int t, int n, const Col& out, const Bq& hp, BlockClock& clk, float* rows_in, size_t stride,
                    bool may_end) {
        if (!af_finite(v)) v = 0.0f;
        if (DC_HP) {
            const float dc = v - x1 + 0.995f * y1;
            x1 = v;
            y1 = dc;
            v = (float)bq_step((double)dc, hp, z1, z2);
            if (!af_finite(v)) v = 0.0f;  // python_api.rs:517-520 runs after the input stage
        }
        const double sq = (double)v * (double)v;
        sum_in += sq;
        blk_in += sq;
        peak_in = fmaxf(peak_in, fabsf(v));
        out.set(t, v);
        if (may_end && clk.at_end(n)) {
            const float rms = (float)sqrt(blk_in / (double)clk.block_len(n));
            rows_in[(size_t)clk.blk * stride] = lin_to_db_f32(rms);
            blk_in = 0.0;
            clk.advance();
        }
    }

    // Register tiles of 8 samples with the next tile's loads issued first: the source is read straight from HBM
    // (a stream's own 32-byte sector per tile when every stream has its own passage, a broadcast line when the
    // streams share one), and with one warp per SM only loads in flight hide that latency.
    template <bool DC_HP>
    AF_HD void run(const float* src, const Col& out, int n0, int len, const Bq& hp, BlockClock clk, float* rows_in,
                   size_t stride) {
        constexpr int U = 8;
        const float* p = src + n0;
        float nxt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) nxt[u] = u < len ? p[u] : 0.0f;
        int t0 = 0;
        for (; t0 + U <= len; t0 += U) {
            float x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = nxt[u];
            if (t0 + 2 * U <= len) {
#pragma unroll
                for (int u = 0; u < U; ++u) nxt[u] = p[t0 + U + u];
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) nxt[u] = t0 + U + u < len ? p[t0 + U + u] : 0.0f;
            }
            const bool may_end = !clk.ends_after(n0 + t0, U);
#pragma unroll
            for (int u = 0; u < U; ++u) step<DC_HP>(x[u], t0 + u, n0 + t0 + u, out, hp, clk, rows_in, stride, may_end);
        }
        for (int u = 0; t0 + u < len; ++u) step<DC_HP>(nxt[u], t0 + u, n0 + t0 + u, out, hp, clk, rows_in, stride, true);
    }
};

// ---- EQ: K consecutive sections of the cascade (dsp/eq.rs:317-322, dsp/biquad.rs) -------------------------
constexpr int kEqTile = 8;
template <int K>
struct EqStage {
    Bq c[K];
    double z[K][2];
    int lane_cnt;  // of the K sections, the ones this stream's candidate really has

    AF_HD void init(const CandidateParams& p, int first_section) {
        const int have = (int)p.n_sections - first_section;
        lane_cnt = have < 0 ? 0 : (have > K ? K : have);
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int s = first_section + j < kMaxSections ? first_section + j : kMaxSections - 1;
            c[j] = bq_from(p.eq[s]);
            z[j][0] = 0.0;
            z[j][1] = 0.0;
        }
    }
    // state only (coefficients are re-read from the candidate every chunk)
    template <class IO>
    AF_HD void sync(IO& io) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            io.f64(z[j][0]);
            io.f64(z[j][1]);
        }
    }

    // samples [t_begin, len) of the chunk, in register tiles of kEqTile samples: the next tile's
    // loads are issued before the current tile is computed (the in-place stores would otherwise pin
    // every load behind them), and the fully unrolled tile x section block of DF2T updates exposes
    // the (sample, section) wavefront parallelism to the instruction scheduler.
    AF_HD void run(const Col& io_chunk, int t_begin, int len) {
        constexpr int U = kEqTile;
        const Col io{io_chunk.base + (size_t)t_begin * io_chunk.stride, io_chunk.stride};
        const int m = len - t_begin;
        if (m <= 0) return;
        float nxt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) nxt[u] = u < m ? io.get(u) : 0.0f;
        for (int t0 = 0; t0 < m; t0 += U) {
            float x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = nxt[u];
            if (t0 + U < m) {
#pragma unroll
                for (int u = 0; u < U; ++u) nxt[u] = t0 + U + u < m ? io.get(t0 + U + u) : 0.0f;
            }
            if (t0 + U <= m) {
#pragma unroll
                for (int j = 0; j < K; ++j) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const float yy = (float)bq_step((double)x[u], c[j], z[j][0], z[j][1]);
                        x[u] = j < lane_cnt ? yy : x[u];
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) io.set(t0 + u, x[u]);
            } else {  // ragged tail: only the real samples may touch the state
#pragma unroll
                for (int j = 0; j < K; ++j) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (t0 + u < m) {
                            const float yy = (float)bq_step((double)x[u], c[j], z[j][0], z[j][1]);
                            x[u] = j < lane_cnt ? yy : x[u];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (t0 + u < m) io.set(t0 + u, x[u]);
            }
        }
    }

    // Head of a legacy render (samples n < F of chunk 0): each section runs the constructor's filter
    // and the configured one side by side and blends them (dsp/eq.rs:279-298 ->
    // dsp/biquad.rs:249-260,290-327).  dflt = constructor coefficients per band ([10][5]).
    // Returns the number of samples handled.
    AF_HD int run_fade_head(const Col& io, int len, int fade_total, int first_section, const double* dflt) {
        const int hlen = len < fade_total ? len : fade_total;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (j >= lane_cnt) continue;
            const int band = first_section + j < 10 ? first_section + j : 9;
            const Bq d = bq_from(dflt + 5 * band);
            double pz1 = 0.0, pz2 = 0.0;
            for (int t = 0; t < hlen; ++t) {
                const double x = (double)io.get(t);
                const double y = bq_step_fading(x, d, c[j], z[j][0], z[j][1], pz1, pz2, t, fade_total);
                io.set(t, (float)y);
            }
            if (hlen == fade_total) {  // promote_pending_coefficients (dsp/biquad.rs:276-286)
                z[j][0] = pz1;
                z[j][1] = pz2;
            }
        }
        return hlen;
    }
};

// ---- dynamic-EQ de-esser (dsp/deesser.rs:405-547) -----------------------------------------------------------
// The five coefficients share the divisor a0 (one refined reciprocal, afsim_math.h); a1 / a0 is the same quotient as
// b1 / a0 (both numerators are -2 cos w).
AF_HD Bq design_peaking(double cs, double alpha, double gain_db) {  // dsp/biquad.rs:110-126,180-181
    const double a = af_exp10(af_div_const(gain_db, 40.0, 0.025));
    const double b0 = 1.0 + alpha * a;
    const double b1 = -2.0 * cs;
    const double b2 = 1.0 - alpha * a;
    const double a0 = 1.0 + alpha / a;
    const double a2 = 1.0 - alpha / a;
    const AfDivisor d = af_divisor(a0);
    Bq c;
    c.b0 = af_div(b0, d);
    c.b1 = af_div(b1, d);
    c.b2 = af_div(b2, d);
    c.a1 = c.b1;
    c.a2 = af_div(a2, d);
    return c;
}

// norm_range(v, S, E) for literal bounds: the divisor E - S and its reciprocal are compile-time constants
// (af_div_const: multiply + exact remainder + correction, bit-identical to the division; afsim_math.h)
#define AF_NORM_RANGE_CONST(v, S, E) clampd(af_div_const((v) - (S), (E) - (S), 1.0 / ((E) - (S))), 0.0, 1.0)
// The parts of the confidence target that do not depend on the band (evaluated once per sample by the map)
struct DeConfShared {
    double voice_conf, narrowness_gain;
};
AF_HD DeConfShared de_confidence_shared(double voice_db, double narrowness) {
    DeConfShared r;
    r.voice_conf = AF_NORM_RANGE_CONST(voice_db, -58.0, -34.0);
    r.narrowness_gain = lerpd(0.35, 1.0, AF_NORM_RANGE_CONST(narrowness, 0.34, 0.68));
    return r;
}
AF_HD double de_confidence_target(double level_db, double voice_db, const DeConfShared& sh) {  // dsp/deesser.rs:171-219
    const double ratio_db = fmax(level_db - voice_db, 0.0);
    const double ratio_conf = AF_NORM_RANGE_CONST(ratio_db, 1.5, 10.0);
    const double level_conf = AF_NORM_RANGE_CONST(level_db, -62.0, -24.0);
    const double narrow_support = (ratio_db > 6.0 && level_db > -45.0) ? 0.75 : 0.0;
    const double voice_support = fmax(sh.voice_conf, narrow_support);
    const double balance_conf = ratio_conf > 0.12 ? fmax(ratio_conf, voice_support * 0.65) : ratio_conf;
    const double broadband_penalty = lerpd(0.35, 1.0, balance_conf);
    return (0.62 * ratio_conf + 0.18 * level_conf + 0.20 * voice_support) * broadband_penalty * sh.narrowness_gain;
}

// Constants of the de-esser in the stream-minor table `base[field * stride]` (base already points
// at this stream's column), so a warp's read of one field is one coalesced transaction.
struct DeConst {
    const double* base;
    size_t stride;
    AF_HD double operator()(int field) const { return base[(size_t)field * stride]; }
    AF_HD Bq bq(int field) const { return bq_from_strided(base + (size_t)field * stride, stride); }
};

// The RMS leg of the blended detector (dsp/compressor.rs:681-686) is db_to_linear(linear_to_db(sqrt(e)))
// = 10^(log10(max(sqrt(e), 1e-10))), i.e. max(sqrt(e), 1e-10) up to the rounding of the log / exp pair
// (~2e-16 relative).  The device map evaluates the closed form -- the same size of deviation as the device
// libm's ulp differences that the parity tolerance already covers -- and saves a log10 + exp10 per sample.
// The host build (tests/hostsim) keeps the reference's round trip so that it stays bit-exact with the oracle.
AF_HD double rms_linear(double env_sq) {
#if defined(__CUDA_ARCH__)
    return fmax(sqrt(env_sq), 1e-10);
#else
    return db_to_lin(lin_to_db(sqrt(env_sq), 1e-10));
#endif
}

// ---- compressor (dsp/compressor.rs:725-774), auto-makeup off ---------------------------------------------
// A chunk is processed in micro-tiles of M samples, phase by phase: the short recurrences (sidechain
// high-pass, band / peak / RMS envelopes, gain-reduction smoothing) run serially, the transcendental
// maps between them (log10 / exp10 / sqrt) are independent across the M samples and overlap.
constexpr int kCompMicro = 4;

struct CompressorStage {
    double threshold, factor, knee, knee_start, knee_end;
    double attack, one_m_attack, det_release, one_m_det_release, release, one_m_release;
    double rms_c, one_m_rms, makeup_lin, sc_c, band_c, one_m_band;
    double fast_c, one_m_fast, charge_c, one_m_charge, slow_c;
    double prev_in, prev_out, low_sq, voiced_sq, presence_sq, peak_env, rms_env, gr, fast_env, slow_env;
    bool adaptive, sidechain;

    AF_HD void init(const CandidateParams& p) {
        threshold = p.c_threshold;
        factor = p.c_factor;
        knee = p.c_knee;
        knee_start = threshold - knee / 2.0;
        knee_end = threshold + knee / 2.0;
        attack = p.c_attack;
        one_m_attack = 1.0 - attack;
        det_release = p.c_det_release;
        one_m_det_release = 1.0 - det_release;
        release = p.c_release;
        one_m_release = 1.0 - release;
        rms_c = p.c_rms;
        one_m_rms = 1.0 - rms_c;
        makeup_lin = p.c_makeup_lin;
        sc_c = p.c_sc;
        band_c = p.c_band;
        one_m_band = 1.0 - band_c;
        fast_c = p.c_fast;
        one_m_fast = 1.0 - fast_c;
        charge_c = p.c_charge;
        one_m_charge = 1.0 - charge_c;
        slow_c = p.c_slow;
        prev_in = prev_out = low_sq = voiced_sq = presence_sq = 0.0;
        peak_env = -120.0;
        rms_env = 0.0;
        gr = fast_env = slow_env = 0.0;
        adaptive = (p.flags & LF_C_ADAPTIVE) != 0;
        sidechain = (p.flags & LF_C_SIDECHAIN) != 0;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
        io.f64(prev_in);
        io.f64(prev_out);
        io.f64(low_sq);
        io.f64(voiced_sq);
        io.f64(presence_sq);
        io.f64(peak_env);
        io.f64(rms_env);
        io.f64(gr);
        io.f64(fast_env);
        io.f64(slow_env);
    }

    AF_HD double gain_computer(double detector_db) const {  // compute_gain_reduction :657-678
        if (knee <= 0.0) return detector_db <= threshold ? 0.0 : (detector_db - threshold) * factor;
        if (detector_db <= knee_start) return 0.0;
        if (detector_db >= knee_end) return (detector_db - threshold) * factor;
        const double x = detector_db - knee_start;
        return factor * x * x / (2.0 * knee);
    }

    // The same walk with the front of the stage (sidechain high-pass, band envelopes, detector weight, instantaneous
    // peak: phases 1 - 2) taken from a shared render: xs / det / wdb / ipk are columns (pitch `ustride`) of the
    // distinct (passage, EQ) pair this stream belongs to.  Phases 3 - 6 are those of `run`, value for value.
    AF_HD void run_shared(const float* xs, const double* dcol, const double* wcol, const double* icol, size_t ustride,
                          const Col& out, int n0, int len, BlockClock clk, float* rows_comp, size_t stride) {
        constexpr int M = kCompMicro;
        for (int t0 = 0; t0 < len; t0 += M) {
            double x[M], det[M], wdb[M], ipk[M], pk[M], rms[M], tgt[M], grv[M];
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const bool live = t0 + i < len;
                const size_t o = (size_t)(t0 + i) * ustride;
                x[i] = live ? (double)xs[o] : 0.0;
                det[i] = live ? dcol[o] : 0.0;
                wdb[i] = live ? wcol[o] : 0.0;
                ipk[i] = live ? icol[o] : lin_to_db(0.0, 1e-10);
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {  // phase 3
                const bool up = ipk[i] > peak_env;
                peak_env = (up ? attack : det_release) * peak_env + (up ? one_m_attack : one_m_det_release) * ipk[i];
                rms_env = rms_c * rms_env + one_m_rms * (det[i] * det[i]);
                pk[i] = peak_env;
                rms[i] = rms_env;
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {  // phase 4
                const double blended = 0.6 * db_to_lin(pk[i]) + 0.4 * rms_linear(rms[i]);
                tgt[i] = gain_computer(lin_to_db(blended, 1e-10) + wdb[i]);
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {  // phase 5
                if (t0 + i < len) {
                    const double target = tgt[i];
                    if (!adaptive) {
                        const bool up = target > gr;
                        gr = (up ? attack : release) * gr + (up ? one_m_attack : one_m_release) * target;
                    } else {
                        if (target > gr)
                            fast_env = attack * gr + one_m_attack * target;
                        else
                            fast_env = fast_c * fast_env + one_m_fast * target;
                        if (target > 3.0)
                            slow_env = charge_c * slow_env + one_m_charge * target;
                        else
                            slow_env *= slow_c;
                        gr = fmax(fast_env, slow_env);
                    }
                }
                grv[i] = gr;
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {  // phase 6
                if (t0 + i < len) {
                    const double gain = db_to_lin(-grv[i]) * makeup_lin;
                    out.set(t0 + i, (float)(x[i] * gain));
                    const int n = n0 + t0 + i;
                    if (clk.at_end(n)) {
                        rows_comp[(size_t)clk.blk * stride] = (float)grv[i];
                        clk.advance();
                    }
                }
            }
        }
    }

    // rows_comp: row.2 table of this stream.  Samples past `len` in the last micro-tile are padding:
    // they touch the recurrence state after the final real sample only, and are never stored.
    AF_HD void run(const Col& io, int n0, int len, BlockClock clk, float* rows_comp, size_t stride) {
        constexpr int M = kCompMicro;
        for (int t0 = 0; t0 < len; t0 += M) {
            double x[M], det[M], wdb[M], ipk[M], pk[M], rms[M], tgt[M], grv[M];
            // phase 1: sidechain high-pass + band envelopes (:407-450), serial
            if (sidechain) {
                double lsq[M], vsq[M], psq[M];
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    x[i] = t0 + i < len ? (double)io.get(t0 + i) : 0.0;
                    const double d = sc_c * (prev_out + x[i] - prev_in);
                    prev_in = x[i];
                    prev_out = d;
                    det[i] = d;
                    const double low = x[i] - d;
                    const double presence = 0.65 * d + 0.35 * (d - low);
                    low_sq = band_c * low_sq + one_m_band * low * low;
                    voiced_sq = band_c * voiced_sq + one_m_band * d * d;
                    presence_sq = band_c * presence_sq + one_m_band * presence * presence;
                    lsq[i] = low_sq;
                    vsq[i] = voiced_sq;
                    psq[i] = presence_sq;
                }
#pragma unroll
                for (int i = 0; i < M; ++i) {  // phase 2: detector weight, independent per sample
                    const double low_rms = sqrt(lsq[i]);
                    const double voiced_rms = fmax(sqrt(vsq[i]), 1e-8);
                    const double presence_rms = sqrt(psq[i]);
                    const AfDivisor by_voiced = af_divisor(voiced_rms);
                    const double plosive = clampd(af_div(low_rms, by_voiced), 0.0, 32.0);
                    const double amount = clampd(af_div_const(plosive - 1.25, 5.0 - 1.25, 1.0 / (5.0 - 1.25)), 0.0, 1.0);
                    const double penalty = 1.0 - amount * (1.0 - 0.35);
                    const double presence_ratio = clampd(af_div(presence_rms, by_voiced), 0.0, 4.0);
                    const double pw = 1.0 + 0.18 * clampd(presence_ratio - 0.75, 0.0, 1.0);
                    wdb[i] = lin_to_db(clampd(penalty * pw, 0.35, 1.15), 1e-10);
                }
            } else {
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    x[i] = t0 + i < len ? (double)io.get(t0 + i) : 0.0;
                    det[i] = x[i];
                    wdb[i] = 0.0;
                }
            }
#pragma unroll
            for (int i = 0; i < M; ++i) ipk[i] = lin_to_db(fabs(det[i]), 1e-10);
            // phase 3: peak (dB domain) and RMS envelopes, serial
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const bool up = ipk[i] > peak_env;
                peak_env = (up ? attack : det_release) * peak_env + (up ? one_m_attack : one_m_det_release) * ipk[i];
                rms_env = rms_c * rms_env + one_m_rms * (det[i] * det[i]);
                pk[i] = peak_env;
                rms[i] = rms_env;
            }
            // phase 4: blended detector (:681-686) + gain computer, independent per sample
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const double blended = 0.6 * db_to_lin(pk[i]) + 0.4 * rms_linear(rms[i]);
                tgt[i] = gain_computer(lin_to_db(blended, 1e-10) + wdb[i]);
            }
            // phase 5: gain-reduction smoothing (:468-505), serial
#pragma unroll
            for (int i = 0; i < M; ++i) {
                if (t0 + i < len) {
                    const double target = tgt[i];
                    if (!adaptive) {
                        const bool up = target > gr;
                        gr = (up ? attack : release) * gr + (up ? one_m_attack : one_m_release) * target;
                    } else {
                        if (target > gr)
                            fast_env = attack * gr + one_m_attack * target;
                        else
                            fast_env = fast_c * fast_env + one_m_fast * target;
                        if (target > 3.0)
                            slow_env = charge_c * slow_env + one_m_charge * target;
                        else
                            slow_env *= slow_c;
                        gr = fmax(fast_env, slow_env);
                    }
                }
                grv[i] = gr;
            }
            // phase 6: apply gain, independent per sample
#pragma unroll
            for (int i = 0; i < M; ++i) {
                if (t0 + i < len) {
                    const double gain = db_to_lin(-grv[i]) * makeup_lin;
                    io.set(t0 + i, (float)(x[i] * gain));
                    const int n = n0 + t0 + i;
                    if (clk.at_end(n)) {
                        rows_comp[(size_t)clk.blk * stride] = (float)grv[i];
                        clk.advance();
                    }
                }
            }
        }
    }
};

// ---- lookahead limiter (dsp/limiter.rs:246-284) ---------------------------------------------------------------
// out[n] = clamp(x[n-L] * g[n]); g follows min(target, release-smoothed target) where the target comes
// from max|x| over [n-L, n].  The reference keeps a monotonic deque; here the sliding maximum over
// W = L+1 samples is the van Herk / Gil-Werman decomposition: time is cut into blocks of W samples,
// the window of sample n is the suffix of the previous block starting at n-L plus the prefix of the
// current block up to n.  Suffix maxima of a finished block are written to a [W][streams] scratch by
// a backward walk at the block's last sample; prefix maxima are a running register.  Exact (max is
// associative and commutative), O(1) amortised per sample for any L.
// NaN: the reference's queue pops every entry a NaN is compared with (`tail_peak > NaN` is false) and the NaN entry is
// itself popped by the next sample, so a NaN at index k removes the samples up to k from the windows of k+1 .. k+L;
// `front.max(|x[n]|)` ignores a NaN operand (dsp/limiter.rs:216-237,253).  fmaxf drops NaN operands but keeps the
// samples before them, so a window that holds a NaN is recomputed by nan_aware_window below (a de-esser / EQ that went
// unstable mid-chain is what feeds NaN to the limiter: the chain's input itself is sanitised).
// x(m): |sample n + m| for m in [-L, 0], 0 before the render's start.
template <class X>
AF_HD float nan_aware_window(const X& x, int L) {
    float w = 0.0f;
    for (int m = -L; m < 0; ++m) {
        const float v = x(m);
        w = v != v ? 0.0f : fmaxf(w, v);
    }
    const float v = x(0);
    return v != v ? w : fmaxf(w, v);
}

struct LimiterStage {
    double g, min_g;
    float prefix;
    uint32_t since_nan;  // samples since the last NaN input (saturating): <= L means the window holds one

    AF_HD void init() {
        g = 1.0;
        min_g = 1.0;
        prefix = 0.0f;
        since_nan = 0x3fffffffu;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
        io.f64(g);
        io.f64(min_g);
        io.f32(prefix);
        io.u32(since_nan);
    }

    // in_ring / out_ring: this stream's column of the whole ring (row r at base[r * stride]);
    // row0 = ring row of chunk sample 0; sfx: this stream's column of the suffix scratch.
    AF_HD void run(const float* in_ring, float* out_ring, size_t stride, int ring_rows, int row0, int n0, int len,
                   int L, double ceil_lin, double rel, float* sfx) {
        const double one_m_rel = 1.0 - rel;
        const int W = L + 1;
        int pos = n0 % W;  // position of sample n inside its block
        for (int t = 0; t < len; ++t) {
            const int n = n0 + t;
            const float in = in_ring[(size_t)(row0 + t) * stride];
            const float in_abs = fabsf(in);
            if (pos == 0) prefix = 0.0f;
            prefix = fmaxf(prefix, in_abs);
            float window = prefix;
            float delayed = 0.0f;
            if (n >= L) {
                int r = row0 + t - L;
                if (r < 0) r += ring_rows;
                delayed = in_ring[(size_t)r * stride];
                // pos == L: the window is exactly the current block (covers n == L, the only n >= L in
                // block 0); otherwise n >= W and the scratch holds the previous block's suffix maxima.
                if (pos != L) window = fmaxf(window, sfx[(size_t)(pos + 1) * stride]);
            }
            since_nan = in != in ? 0u : (since_nan < 0x3fffffffu ? since_nan + 1u : since_nan);
            if (since_nan <= (uint32_t)L) {  // rare: the exact queue semantics around a NaN
                auto x = [&](int m) {
                    if (n + m < 0) return 0.0f;
                    int r = row0 + t + m;
                    if (r < 0) r += ring_rows;
                    return fabsf(in_ring[(size_t)r * stride]);
                };
                window = nan_aware_window(x, L);
            }
            const double peak = (double)window;
            const double target = peak > ceil_lin ? ceil_lin / peak : 1.0;
            if (target < g)
                g = target;
            else
                g = rel * g + one_m_rel * target;
            min_g = fmin(min_g, g);
            out_ring[(size_t)(row0 + t) * stride] = (float)clampd((double)delayed * g, -ceil_lin, ceil_lin);
            if (pos == L) {  // block finished: suffix maxima for the next block's windows
                float m = 0.0f;
                for (int i = L; i >= 1; --i) {
                    int r = row0 + t - (L - i);
                    if (r < 0) r += ring_rows;
                    m = fmaxf(m, fabsf(in_ring[(size_t)r * stride]));
                    sfx[(size_t)i * stride] = m;
                }
                pos = 0;
            } else {
                pos += 1;
            }
        }
    }

    AF_HD float peak_reduction_db() const {  // dsp/limiter.rs:272-279; monotone in g, so one conversion of min g
        return min_g < 1.0 ? (float)(-lin_to_db(min_g, 1e-10)) : 0.0f;
    }
};

// ---- 4x polyphase true-peak FIR (dsp/true_peak.rs:173-186) --------------------------------------------------
// 8 consecutive outputs x 4 phases from a register window; taps accumulated k = 0..31 in the
// reference's order with fused multiply-add.  win[i] = x[first - 31 + i], i = 0..38.
constexpr int kFirChunk = 8;
constexpr int kFirWin = 31 + kFirChunk;

template <typename FIR>
AF_HD void fir8_peaks(const float (&win)[kFirWin], const FIR& fir, float (&peak)[kFirChunk]) {
    float acc[kFirChunk][4];
#pragma unroll
    for (int j = 0; j < kFirChunk; ++j)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[j][p] = 0.0f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float c = fir[p][k];
#pragma unroll
            for (int j = 0; j < kFirChunk; ++j) acc[j][p] = af_fmaf(c, win[31 + j - k], acc[j][p]);
        }
    }
#pragma unroll
    for (int j = 0; j < kFirChunk; ++j) {
        float pk = fabsf(win[31 + j]);
#pragma unroll
        for (int p = 0; p < 4; ++p) pk = fmaxf(pk, fabsf(acc[j][p]));
        peak[j] = pk;
    }
}

// ---- true-peak limiter (dsp/true_peak.rs:337-378) ------------------------------------------------------------
struct TpLimiterStage {
    float win[kFirWin];  // win[0..30] = the last 31 (sanitized) inputs; the 20-sample delay reads from it
    float g, min_g, peak_pre;
    uint32_t events;
    bool limited;

    AF_HD void init() {
#pragma unroll
        for (int i = 0; i < kFirWin; ++i) win[i] = 0.0f;
        g = 1.0f;
        min_g = 1.0f;
        peak_pre = 0.0f;
        events = 0;
        limited = false;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
#pragma unroll
        for (int i = 0; i < 31; ++i) io.f32(win[i]);
        io.f32(g);
        io.f32(min_g);
        io.f32(peak_pre);
        io.u32(events);
        io.flag(limited);
    }

    // One group of kFirChunk samples: in[j] -> out[j]; `valid` of them are real.
    template <typename FIR>
    AF_HD void group(const float (&in)[kFirChunk], float (&out)[kFirChunk], int valid, int n_first, BlockClock& clk,
                     const FIR& fir, float ceil_lin, float rel, float one_m_rel) {
#pragma unroll
        for (int j = 0; j < kFirChunk; ++j) win[31 + j] = af_finite(in[j]) ? in[j] : 0.0f;
        float itp[kFirChunk];
        fir8_peaks(win, fir, itp);
#pragma unroll
        for (int j = 0; j < kFirChunk; ++j) {
            out[j] = 0.0f;
            if (j < valid) {
                peak_pre = fmaxf(peak_pre, itp[j]);
                const float target = itp[j] > ceil_lin ? clampf((ceil_lin * 0.999f) / itp[j], 0.0f, 1.0f) : 1.0f;
                if (target < g) {
                    g = target;
                    limited = true;
                } else {
                    g = rel * g + one_m_rel * target;
                }
                min_g = fminf(min_g, g);
                float o = clampf(win[31 + j - kTpDelay] * g, -ceil_lin, ceil_lin);
                if (!af_finite(o)) o = 0.0f;
                out[j] = o;
                if (clk.at_end(n_first + j)) {
                    events += limited ? 1u : 0u;
                    limited = false;
                    clk.advance();
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 31; ++i) win[i] = win[i + kFirChunk];
    }

    AF_HD float peak_reduction_db() const {  // dsp/true_peak.rs:320-326, monotone in g
        return min_g >= 1.0f ? 0.0f : -20.0f * af_log10_f32(fmaxf(min_g, 1e-10f));
    }
};

// ---- true-peak detector + output statistics (dsp/true_peak.rs:208-218, python_api.rs:529-575) ----------------
struct OutputStage {
    float win[kFirWin];
    double sum_out, blk_out;
    float peak_out, peak_tp;
    bool non_finite;

    AF_HD void init() {
#pragma unroll
        for (int i = 0; i < kFirWin; ++i) win[i] = 0.0f;
        sum_out = 0.0;
        blk_out = 0.0;
        peak_out = 0.0f;
        peak_tp = 0.0f;
        non_finite = false;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
#pragma unroll
        for (int i = 0; i < 31; ++i) io.f32(win[i]);
        io.f64(sum_out);
        io.f64(blk_out);
        io.f32(peak_out);
        io.f32(peak_tp);
        io.flag(non_finite);
    }

    // rows_out: row.1 table of this stream (nullptr: detector only)
    template <typename FIR>
    AF_HD void group(const float (&v)[kFirChunk], int valid, int n_first, BlockClock& clk, const FIR& fir,
                     float* rows_out, size_t stride) {
#pragma unroll
        for (int j = 0; j < kFirChunk; ++j) win[31 + j] = af_finite(v[j]) ? v[j] : 0.0f;
        float otp[kFirChunk];
        fir8_peaks(win, fir, otp);
#pragma unroll
        for (int j = 0; j < kFirChunk; ++j) {
            if (j < valid) {
                const float s = v[j];
                peak_out = fmaxf(peak_out, fabsf(s));
                peak_tp = fmaxf(peak_tp, otp[j]);
                const double sq = (double)s * (double)s;
                sum_out += sq;
                if (af_finite(s))
                    blk_out += sq;
                else
                    non_finite = true;
                if (clk.at_end(n_first + j)) {
                    if (rows_out) {
                        const float rms = (float)sqrt(blk_out / (double)clk.block_len(n_first + j));
                        rows_out[(size_t)clk.blk * stride] = lin_to_db_f32(rms);
                    }
                    blk_out = 0.0;
                    clk.advance();
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 31; ++i) win[i] = win[i + kFirChunk];
    }
};

}  // namespace afsim
