// afsim_split.h -- the chain stages split into thin serial recurrences (R) and parallel maps (M).
//
// Few-stream sweeps (4096 candidates x one 30 s passage is only 128 warps) cannot fill a B200 with
// one thread per stream: a warp that walks a recurrence is latency bound at ~0.25 instructions per
// cycle.  But almost all of the chain's arithmetic is feed-forward: the transcendental maps of the
// compressor (log10 / exp10 / sqrt), the limiter's window maximum and target gain, and both 4x
// polyphase true-peak FIRs depend on the recurrences only through a few per-sample values.  So every
// stage is cut along that line:
//     R kernels   one thread per stream walks the chunk (short dependent chain, a few flops per sample)
//     M kernels   one thread per (stream, 8 samples): fully parallel over time, fills every SM
// The hand-off arrays are time-major rings ([row][stream], f64 `w0..w3` and f32 `buf_*`), so both
// kinds of kernel access them coalesced.  Per sample the operations and their order are exactly those
// of the fused stage in afsim_stages.h (and of the reference), so results are bit-identical to the
// fused path; tests/hostsim checks that on the CPU.
//
// Reference citations are relative to rust-core/src/.
#pragma once
#if defined(__CUDACC__)
#include <cuda_pipeline.h>
#endif
#include "afsim_stages.h"

namespace afsim {

constexpr int kGroup = 8;         // samples per tile of the serial kernels / per FIR and limiter map thread
constexpr int kCompMapGroup = 2;  // samples per thread of the compressor maps: their transcendental chains need
                                  // ~100+ registers per sample in flight, so occupancy (not unrolling) hides latency

AF_HD int ring_row(int row0, int t, int ring_rows) {
    int r = row0 + t;
    if (r < 0) r += ring_rows;
    return r;
}

// ---- asynchronous staging of the serial kernels' inputs --------------------------------------------------------
// A thread that walks a recurrence reads a few bytes per sample from hand-off rings that live in HBM
// (a chunk of every ring is tens of MB, far more than L2 holds across the ~25 stages in flight), and
// with one warp per SM nothing hides that ~1 us latency.  So each R thread streams its inputs through a
// private ring in shared memory with cp.async (LDGSTS): kPipeDepth tiles of kGroup samples are in
// flight while the current tile is walked.  On the host the copies complete immediately, so
// tests/hostsim exercises the same pipeline logic.
constexpr int kPipeDepth = 8;

AF_HD void async_copy(float* dst, const float* src) {
#if defined(__CUDA_ARCH__)
    __pipeline_memcpy_async(dst, src, 4);
#else
    *dst = *src;
#endif
}
AF_HD void async_copy(double* dst, const double* src) {
#if defined(__CUDA_ARCH__)
    __pipeline_memcpy_async(dst, src, 8);
#else
    *dst = *src;
#endif
}
AF_HD void async_commit() {
#if defined(__CUDA_ARCH__)
    __pipeline_commit();
#endif
}
template <int N>
AF_HD void async_wait_prior() {
#if defined(__CUDA_ARCH__)
    __pipeline_wait_prior(N);
#endif
}

template <typename T, int DEPTH = kPipeDepth>
struct StageRing {  // one staged input stream of one thread
    T* p;           // this thread's element of (tile slot 0, sample 0); nullptr = direct mode (no staging)
    int lanes;      // threads sharing the staging area (element pitch)
    bool on;        // staged (a compile-time constant in the kernels: the accessors below fold to one path)
    AF_HD T* at(int tile, int u) const { return p + (size_t)((tile % DEPTH) * kGroup + u) * lanes; }
    // direct-mode aware accessors: big batches have enough warps per SM to hide the latency themselves, and
    // the staging area would only cost them occupancy
    AF_HD void fetch(int tile, int u, const T* g) const {
        if (on) async_copy(at(tile, u), g);
    }
    AF_HD T get(int tile, int u, const T* g) const { return on ? *at(tile, u) : *g; }
};
struct Staging {  // the staging area of a block (shared memory) / of one call (host); base == nullptr: direct mode
    unsigned char* base;
    int lanes, lane;
    size_t used;
    // 1 / 0: staged / direct, spelled out by the kernels so that it is a compile-time constant there (a shared-memory
    // address cannot be proven non-null: every staged access would carry a pointer test and both code paths);
    // -1: decided by `base` at run time (the host harness)
    int mode = -1;
    AF_HD bool staged() const { return mode < 0 ? base != nullptr : mode != 0; }
    template <typename T, int DEPTH = kPipeDepth>
    AF_HD StageRing<T, DEPTH> ring() {
        StageRing<T, DEPTH> r;
        r.on = staged();
        r.p = r.on ? reinterpret_cast<T*>(base + used) + lane : nullptr;
        r.lanes = lanes;
        used += sizeof(T) * (size_t)DEPTH * kGroup * (size_t)lanes;
        return r;
    }
};
constexpr size_t kStagingBytesPerLane = (size_t)kPipeDepth * 8 * 16;  // widest user: two f64 streams (kGroup = 8)

// issue(k, full) queues the copies of tile k; body(k, full) runs once tile k has landed.  `full` is a
// compile-time flag (TileFull / TileRagged): all tiles but the last one of a render hold kGroup samples, and
// their code carries no per-sample bounds predicates.
struct TileFull { static constexpr bool value = true; };
struct TileRagged { static constexpr bool value = false; };
template <int DEPTH, class Issue, class Body>
AF_HD void pipelined_tiles_depth(int len, Issue issue, Body body) {
    const int n_tiles = (len + 8 - 1) / 8;
    const int n_full = len / 8;
    for (int k = 0; k < DEPTH - 1; ++k) {
        if (k < n_full)
            issue(k, TileFull());
        else if (k < n_tiles)
            issue(k, TileRagged());
        async_commit();
    }
    for (int k = 0; k < n_tiles; ++k) {
        const int ahead = k + DEPTH - 1;
        if (ahead < n_full)
            issue(ahead, TileFull());
        else if (ahead < n_tiles)
            issue(ahead, TileRagged());
        async_commit();
        async_wait_prior<DEPTH - 1>();
        if (k < n_full)
            body(k, TileFull());
        else
            body(k, TileRagged());
    }
}
template <class Issue, class Body>
AF_HD void pipelined_tiles(int len, Issue issue, Body body) {
    pipelined_tiles_depth<kPipeDepth>(len, issue, body);
}

template <typename T, int U>
AF_HD void load_tile(const T* col, size_t stride, int valid, T (&v)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = u < valid ? col[(size_t)u * stride] : (T)0;
}
template <typename T, int U>
AF_HD void store_tile(T* col, size_t stride, int valid, const T (&v)[U]) {
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (u < valid) col[(size_t)u * stride] = v[u];
}

// ---- compressor (dsp/compressor.rs:725-774) ------------------------------------------------------------------
// Constants come from CompressorStage::init; the recurrence state of the three serial passes lives in
// the same slots of the compressor state table as the fused stage uses.
struct CompSplit : CompressorStage {
    template <class IO>
    AF_HD void sync_r1(IO& io) {
        io.f64(prev_in);
        io.f64(prev_out);
        io.f64(low_sq);
        io.f64(voiced_sq);
        io.f64(presence_sq);
    }
    template <class IO>
    AF_HD void sync_r3(IO& io) {  // table offset 5
        io.f64(peak_env);
        io.f64(rms_env);
    }
    template <class IO>
    AF_HD void sync_r5(IO& io) {  // table offset 7
        io.f64(gr);
        io.f64(fast_env);
        io.f64(slow_env);
    }

    // R1: sidechain high-pass + band envelopes (:407-450).  x -> w0 = det, w1..w3 = band envelopes^2
    AF_HD void run_r1(const float* x, double* w0, double* w1, double* w2, double* w3, size_t stride, int len, Staging stg) {
        constexpr int U = kGroup;
        const StageRing<float> sx = stg.ring<float>();
        auto issue = [&](int k, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = k * U;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (FULL || t0 + u < len) async_copy(sx.at(k, u), x + (size_t)(t0 + u) * stride);
        };
        auto body = [&](int k, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = k * U;
            const int valid = FULL ? U : len - t0;
            float xin[U];
#pragma unroll
            for (int u = 0; u < U; ++u) xin[u] = (FULL || u < valid) ? *sx.at(k, u) : 0.0f;
            double det[U], lsq[U], vsq[U], psq[U];
            if (sidechain) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (FULL || u < valid) {
                        const double xv = (double)xin[u];
                        const double d = sc_c * (prev_out + xv - prev_in);
                        prev_in = xv;
                        prev_out = d;
                        const double low = xv - d;
                        const double presence = 0.65 * d + 0.35 * (d - low);
                        low_sq = band_c * low_sq + one_m_band * low * low;
                        voiced_sq = band_c * voiced_sq + one_m_band * d * d;
                        presence_sq = band_c * presence_sq + one_m_band * presence * presence;
                        det[u] = d;
                    } else {
                        det[u] = 0.0;
                    }
                    lsq[u] = low_sq;
                    vsq[u] = voiced_sq;
                    psq[u] = presence_sq;
                }
                store_tile(w1 + (size_t)t0 * stride, stride, valid, lsq);
                store_tile(w2 + (size_t)t0 * stride, stride, valid, vsq);
                store_tile(w3 + (size_t)t0 * stride, stride, valid, psq);
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) det[u] = (double)xin[u];
            }
            store_tile(w0 + (size_t)t0 * stride, stride, valid, det);
        };
        pipelined_tiles(len, issue, body);
    }

    // The constants the maps use (gain computer, makeup, sidechain flag) from the batch's stream-minor table
    // (`tab` points at the stream's column of BatchArgs::map_tab); the other members stay unset.
    AF_HD void init_map(const double* tab, size_t stride) {
        threshold = tab[MT_THRESHOLD * stride];
        factor = tab[MT_FACTOR * stride];
        knee = tab[MT_KNEE * stride];
        knee_start = threshold - knee / 2.0;
        knee_end = threshold + knee / 2.0;
        makeup_lin = tab[MT_MAKEUP_LIN * stride];
        sidechain = (static_cast<uint32_t>(tab[MT_FLAGS * stride]) & LF_C_SIDECHAIN) != 0;
    }

    // M2: detector weight in dB and instantaneous peak in dB.  w1 <- wdb, w2 <- ipk
    AF_HD void map_m2(const double* w0, double* w1, double* w2, const double* w3, size_t stride, int valid) const {
        double det[kCompMapGroup], lsq[kCompMapGroup], vsq[kCompMapGroup], psq[kCompMapGroup], wdb[kCompMapGroup], ipk[kCompMapGroup];
        load_tile(w0, stride, valid, det);
        if (sidechain) {
            load_tile((const double*)w1, stride, valid, lsq);
            load_tile((const double*)w2, stride, valid, vsq);
            load_tile(w3, stride, valid, psq);
        }
#pragma unroll
        for (int u = 0; u < kCompMapGroup; ++u) {
            if (sidechain) {
                const double low_rms = sqrt(lsq[u]);
                const double voiced_rms = fmax(sqrt(vsq[u]), 1e-8);
                const double presence_rms = sqrt(psq[u]);
                const AfDivisor by_voiced = af_divisor(voiced_rms);  // one refined reciprocal for both ratios
                const double plosive = clampd(af_div(low_rms, by_voiced), 0.0, 32.0);
                const double amount = clampd(af_div_const(plosive - 1.25, 5.0 - 1.25, 1.0 / (5.0 - 1.25)), 0.0, 1.0);
                const double penalty = 1.0 - amount * (1.0 - 0.35);
                const double presence_ratio = clampd(af_div(presence_rms, by_voiced), 0.0, 4.0);
                const double pw = 1.0 + 0.18 * clampd(presence_ratio - 0.75, 0.0, 1.0);
                wdb[u] = lin_to_db(clampd(penalty * pw, 0.35, 1.15), 1e-10);
            } else {
                wdb[u] = 0.0;
            }
            ipk[u] = lin_to_db(fabs(det[u]), 1e-10);
        }
        store_tile(w1, stride, valid, wdb);
        store_tile(w2, stride, valid, ipk);
    }

    // R3: peak (dB domain) and RMS envelopes.  (det, ipk: the stream's w0 / w2 columns, or the shared compressor front
    // of its (passage, EQ) pair, pitch in_stride) -> w2 = peak_env, w3 = rms_env
    AF_HD void run_r3(const double* det_in, const double* ipk_in, size_t in_stride, double* w2, double* w3, size_t stride, int len,
                      Staging stg) {
        constexpr int U = kGroup;
        const StageRing<double> sdet = stg.ring<double>();
        const StageRing<double> sipk = stg.ring<double>();
        auto issue = [&](int k, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = k * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || t0 + u < len) {
                    async_copy(sdet.at(k, u), det_in + (size_t)(t0 + u) * in_stride);
                    async_copy(sipk.at(k, u), ipk_in + (size_t)(t0 + u) * in_stride);
                }
            }
        };
        auto body = [&](int k, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = k * U;
            const int valid = FULL ? U : len - t0;
            double det[U], ipk[U], pk[U], rms[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                det[u] = (FULL || u < valid) ? *sdet.at(k, u) : 0.0;
                ipk[u] = (FULL || u < valid) ? *sipk.at(k, u) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || u < valid) {
                    const bool up = ipk[u] > peak_env;
                    peak_env = (up ? attack : det_release) * peak_env + (up ? one_m_attack : one_m_det_release) * ipk[u];
                    rms_env = rms_c * rms_env + one_m_rms * (det[u] * det[u]);
                }
                pk[u] = peak_env;
                rms[u] = rms_env;
            }
            store_tile(w2 + (size_t)t0 * stride, stride, valid, pk);
            store_tile(w3 + (size_t)t0 * stride, stride, valid, rms);
        };
        pipelined_tiles(len, issue, body);
    }

    // M4: blended detector (:681-686) + gain computer (:657-678).  (w2 pk, w3 rms, wdb) -> w1 = target GR
    // (wdb_in: the stream's w1 column or the shared front's, pitch wdb_stride)
    AF_HD void map_m4(const double* wdb_in, size_t wdb_stride, double* w1, const double* w2, const double* w3, size_t stride,
                      int valid) const {
        double pk[kCompMapGroup], rms[kCompMapGroup], wdb[kCompMapGroup], tgt[kCompMapGroup];
        load_tile(w2, stride, valid, pk);
        load_tile(w3, stride, valid, rms);
        load_tile(wdb_in, wdb_stride, valid, wdb);
#pragma unroll
        for (int u = 0; u < kCompMapGroup; ++u) {
            tgt[u] = gain_computer(lin_to_db(0.6 * db_to_lin(pk[u]) + 0.4 * rms_linear(rms[u]), 1e-10) + wdb[u]);
        }
        store_tile(w1, stride, valid, tgt);
    }

    // R5: gain-reduction smoothing (:468-505), in place on w1; block-end meter rows
    AF_HD void run_r5(double* w1, size_t stride, int n0, int len, BlockClock clk, float* rows_comp, Staging stg) {
        constexpr int U = kGroup;
        const StageRing<double> stgt = stg.ring<double>();
        auto issue = [&](int k, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = k * U;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (FULL || t0 + u < len) async_copy(stgt.at(k, u), (const double*)w1 + (size_t)(t0 + u) * stride);
        };
        auto body = [&](int k, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = k * U;
            const int valid = FULL ? U : len - t0;
            double tgt[U], grv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) tgt[u] = (FULL || u < valid) ? *stgt.at(k, u) : 0.0;
            auto walk = [&](auto may_end) {  // may_end: an analysis block can end inside this tile
                constexpr bool CHECK = decltype(may_end)::value;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (FULL || u < valid) {
                        const double target = tgt[u];
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
                        if (CHECK && clk.at_end(n0 + t0 + u)) {
                            rows_comp[(size_t)clk.blk * stride] = (float)gr;
                            clk.advance();
                        }
                    }
                    grv[u] = gr;
                }
            };
            if (FULL && clk.ends_after(n0 + t0, U))
                walk(TileRagged());
            else
                walk(TileFull());
            store_tile(w1 + (size_t)t0 * stride, stride, valid, grv);
        };
        pipelined_tiles(len, issue, body);
    }

    // M6: apply gain.  (w1 gr, x_in: the stream's own column or the shared EQ output, pitch x_stride) -> x
    AF_HD void map_m6(const double* w1, const float* x_in, size_t x_stride, float* x, size_t stride, int valid) const {
        double grv[kCompMapGroup];
        float xin[kCompMapGroup], y[kCompMapGroup];
        load_tile(w1, stride, valid, grv);
        load_tile(x_in, x_stride, valid, xin);
#pragma unroll
        for (int u = 0; u < kCompMapGroup; ++u) {
            const double gain = db_to_lin(-grv[u]) * makeup_lin;
            y[u] = (float)((double)xin[u] * gain);
        }
        store_tile(x, stride, valid, y);
    }
    // M6 of an auto-makeup batch: only the gain-reduction factor; the makeup stage R7 applies it.  w1 gr -> w1 = 10^(-gr/20)
    AF_HD void map_m6_gain(double* w1, size_t stride, int valid) const {
        double grv[kCompMapGroup], g[kCompMapGroup];
        load_tile((const double*)w1, stride, valid, grv);
#pragma unroll
        for (int u = 0; u < kCompMapGroup; ++u) g[u] = db_to_lin(-grv[u]);
        store_tile(w1, stride, valid, g);
    }
};

// ---- R7: auto makeup (dsp/compressor.rs:598-653,700-722) over the momentary loudness meter (dsp/loudness.rs) --------
// The makeup gain is constant inside a block and steps at block ends, where the compressor (:700-722)
//   1. estimates speech activity from the block's PRE-gain RMS (and the control simulator's VAD / noise evidence),
//   2. feeds the POST-gain block to the loudness meter if the block was active,
//   3. moves the smoothed makeup towards target_lufs - momentary loudness (update_auto_makeup_gain).
// One thread walks its stream: out = (f32)(x * (g * makeup_lin)), the K-weighting section runs on `out`
// speculatively (whether the block is fed is only known at its end; an unfed block rolls the four filter states
// back), and the 400 ms window is a ring of per-slot partial sums (MakeupConst).  Chunks of an auto-makeup batch are
// whole blocks, so a block never straddles two launches.
AF_HD double af_pow(double a, double b) { return pow(a, b); }

struct MakeupR {
    // state
    double mk, score, rel, lufs;     // smoothed makeup (dB), speech activity score, activity reliability, current_lufs
    double c1, c2, c3, c4;           // meter filter state after the last block that was fed
    uint32_t pos;                    // ring slot where the next fed block starts
    // constants
    double makeup_db, target_lufs, k_smooth, k_relax, k_activity;
    double ev_vad_rel, ev_floor, ev_live_rel, ev_cfg_rel;
    bool evidence;

    AF_HD void init(const CandidateParams& p) {
        makeup_db = p.c_makeup_db;
        target_lufs = p.c_target_lufs;
        k_smooth = p.c_mk_smooth;
        k_relax = p.c_mk_relax;
        k_activity = p.c_mk_activity;
        ev_vad_rel = p.c_ev_vad_reliability;
        ev_floor = p.c_ev_noise_floor_db;
        ev_live_rel = p.c_ev_live_reliability;
        ev_cfg_rel = p.c_ev_cfg_reliability;
        evidence = (p.flags & LF_C_EVIDENCE) != 0;
        mk = p.c_makeup_db;
        score = 0.0;
        rel = 0.0;
        lufs = -100.0;
        c1 = c2 = c3 = c4 = 0.0;
        pos = 0;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
        io.f64(mk);
        io.f64(score);
        io.f64(rel);
        io.f64(lufs);
        io.f64(c1);
        io.f64(c2);
        io.f64(c3);
        io.f64(c4);
        io.u32(pos);
    }

    static AF_HD double activity_from_rms_db(double rms_db) {  // :507-514
        if (!(rms_db >= -55.0 && rms_db <= -6.0)) return 0.0;
        const double onset = clampd((rms_db - -55.0) / 12.0, 0.0, 1.0);
        const double overload = clampd((-6.0 - rms_db) / 6.0, 0.0, 1.0);
        return fmin(onset, overload);
    }
    static AF_HD bool finite_d(double v) { return fabs(v) <= 1.7976931348623157e308; }
    static AF_HD double unit_or_zero(double v, bool* ok) {  // finite_unit, :516-518
        *ok = finite_d(v);
        return *ok ? clampd(v, 0.0, 1.0) : 0.0;
    }
    static AF_HD double smoothstep(double e0, double e1, double v) {  // :520-526
        if (!finite_d(v) || !finite_d(e0) || !finite_d(e1) || e1 <= e0) return 0.0;
        const double t = clampd((v - e0) / (e1 - e0), 0.0, 1.0);
        return t * t * (3.0 - 2.0 * t);
    }
    // estimate_auto_makeup_activity (:528-581); have_vad: this block has a VAD probability
    AF_HD void estimate(double rms_db, bool have_vad, double vad_p_in, double* activity, double* reliability) const {
        const double absolute = activity_from_rms_db(rms_db);
        if (!have_vad) {
            *activity = absolute;
            *reliability = 1.0;
            return;
        }
        bool ok;
        double vad_rel = unit_or_zero(ev_vad_rel, &ok);
        double vad_p = unit_or_zero(vad_p_in, &ok);
        if (!ok) vad_rel = 0.0;
        const double cfg_rel = unit_or_zero(ev_cfg_rel, &ok);
        const double live_rel = unit_or_zero(ev_live_rel, &ok);
        double noise_rel = cfg_rel > 0.0 ? fmin(live_rel, cfg_rel) : live_rel;
        double relative = 0.0;
        if (finite_d(ev_floor) && ev_floor >= -120.0 && ev_floor <= 0.0)
            relative = smoothstep(ev_floor + 3.0, ev_floor + 15.0, rms_db);
        else
            noise_rel = 0.0;
        const double fallback = noise_rel * relative + (1.0 - noise_rel) * absolute;
        *activity = clampd(vad_rel * vad_p + (1.0 - vad_rel) * fallback, 0.0, 1.0);
        *reliability = clampd(fmax(vad_rel, 0.75 * noise_rel), 0.0, 1.0);
    }
    // update_auto_makeup_gain (:598-653) with the switch on and a meter present; limiter feedback is 0 offline
    AF_HD void update(double activity, double reliability, int elapsed) {
        const double n = (double)(elapsed > 1 ? elapsed : 1);
        const double makeup_coeff = af_pow(k_smooth, n);
        const double relax_coeff = af_pow(k_relax, n);
        const double activity_coeff = af_pow(k_activity, n);
        score = activity_coeff * score + (1.0 - activity_coeff) * clampd(activity, 0.0, 1.0);
        rel = clampd(reliability, 0.0, 1.0);
        if (score < 0.20) {
            mk = relax_coeff * mk + (1.0 - relax_coeff) * makeup_db;
            return;
        }
        if (rel < 0.35) {
            const double cap = makeup_db + 3.0 * (rel / 0.35);
            if (mk > cap) mk = makeup_coeff * mk + (1.0 - makeup_coeff) * cap;
            return;
        }
        const double required = target_lufs - lufs;
        const double reliability_cap = clampd(12.0 * rel, 3.0, 12.0);
        const double headroom_cap = clampd(12.0 - 0.0 * 2.0, 0.0, reliability_cap);
        const double clamped = clampd(required, 0.0, headroom_cap);
        const double diff = clamped - mk;
        if (fabs(diff) > 0.1)
            mk = makeup_coeff * mk + (1.0 - makeup_coeff) * clamped;
        else
            mk = clamped;
    }

    // x: compressor input / output column at chunk start (in place); g: 10^(-gr/20) column; ring: this stream's column
    // of [2 * n_slots + 2 * kMaxMakeupSub][stride] (full sums, tail sums, pending full, pending tail); rows: this
    // stream's column of [3][n_rows][stride] or nullptr; vad: this stream's probabilities or nullptr
    AF_HD void run(float* x, const double* g, size_t stride, int n0, int len, BlockClock clk, const MakeupConst& mc, double* ring,
                   float* rows, int n_rows, const double* vad, Staging stg) {
        constexpr int U = kGroup;
        const StageRing<float> sx = stg.ring<float>();
        const StageRing<double> sg = stg.ring<double>();
        const double b0 = mc.b[0], b1 = mc.b[1], b2 = mc.b[2], b3 = mc.b[3], b4 = mc.b[4];
        const double a1 = mc.a[1], a2 = mc.a[2], a3 = mc.a[3], a4 = mc.a[4];
        const int slot = mc.slot, n_slots = mc.n_slots, tail_from = mc.tail_from;
        double* ring_full = ring;
        double* ring_tail = ring + (size_t)n_slots * stride;
        double* pend_full = ring + (size_t)2 * n_slots * stride;
        double* pend_tail = pend_full + (size_t)kMaxMakeupSub * stride;
        double v1 = c1, v2 = c2, v3 = c3, v4 = c4;
        double mk_lin = db_to_lin(mk);
        double sq_in = 0.0, full = 0.0, tail = 0.0;
        int off = 0, sub = 0;  // samples into the current slot, slots of the current block already closed
        auto issue = [&](int k, auto fullt) {
            constexpr bool FULL = decltype(fullt)::value;
            const int t0 = k * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || t0 + u < len) {
                    sx.fetch(k, u, x + (size_t)(t0 + u) * stride);
                    sg.fetch(k, u, g + (size_t)(t0 + u) * stride);
                }
            }
        };
        auto body = [&](int k, auto fullt) {
            constexpr bool FULL = decltype(fullt)::value;
            const int t0 = k * U;
            const int valid = FULL ? U : len - t0;
            float xin[U], y[U];
            double gv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                xin[u] = (FULL || u < valid) ? sx.get(k, u, x + (size_t)(t0 + u) * stride) : 0.0f;
                gv[u] = (FULL || u < valid) ? sg.get(k, u, g + (size_t)(t0 + u) * stride) : 0.0;
            }
            // one sample: gain, output, pre-gain energy, K-weighting (direct form II, add_frames_f32), slot sums
            auto step = [&](int u) {
                const double xv = (double)xin[u];
                const double gain = gv[u] * mk_lin;
                const float o = (float)(xv * gain);
                y[u] = o;
                sq_in += xv * xv;
                const double v0 = (double)o - a1 * v1 - a2 * v2 - a3 * v3 - a4 * v4;
                const double w = b0 * v0 + b1 * v1 + b2 * v2 + b3 * v3 + b4 * v4;
                v4 = v3;
                v3 = v2;
                v2 = v1;
                v1 = fabs(v0) < 2.2250738585072014e-308 ? 0.0 : v0;  // denormal flush of the filter state
                const double w2 = w * w;
                full += w2;
                tail += off >= tail_from ? w2 : 0.0;
                off += 1;
            };
            if (FULL && slot - off > U && clk.ends_after(n0 + t0, U)) {
                // no slot end, no block end in this tile: no per-sample checks
#pragma unroll
                for (int u = 0; u < U; ++u) step(u);
            } else {
                // a slot (and with it possibly the block) can end inside this tile: rolled, so that the block-end code
                // exists once
#pragma unroll 1
                for (int u = 0; u < U; ++u) {
                    y[u] = 0.0f;
                    if (!(FULL || u < valid)) continue;
                    step(u);
                    const int n = n0 + t0 + u;
                    if (off == slot) {
                        pend_full[(size_t)sub * stride] = full;
                        pend_tail[(size_t)sub * stride] = tail;
                        sub += 1;
                        full = tail = 0.0;
                        off = 0;
                    }
                    if (clk.at_end(n)) {
                        const int blen = clk.block_len(n);
                        const double rms_db = lin_to_db(sqrt(sq_in / (double)blen), 1e-10);  // block_rms_db :583-596
                        double activity, reliability;
                        const bool have_vad = evidence && vad != nullptr;
                        estimate(rms_db, have_vad, have_vad ? vad[clk.blk] : 0.0, &activity, &reliability);
                        if (activity > 0.20 && reliability >= 0.35) {  // :713-718: the meter hears the block
                            c1 = v1;
                            c2 = v2;
                            c3 = v3;
                            c4 = v4;
                            for (int i = 0; i < sub; ++i) {
                                const size_t r = (size_t)((pos + i) % n_slots) * stride;
                                ring_full[r] = pend_full[(size_t)i * stride];
                                ring_tail[r] = pend_tail[(size_t)i * stride];
                            }
                            const int partial = off > 0 ? (int)((pos + sub) % n_slots) : -1;  // final short block only
                            double sum = 0.0;
                            for (int i = 0; i < n_slots; ++i)
                                sum += i == partial ? full + ring_tail[(size_t)i * stride] : ring_full[(size_t)i * stride];
                            pos = (pos + sub) % n_slots;
                            const double energy = sum / (double)mc.window;
                            const double l = energy <= 0.0 ? -(double)INFINITY : 10.0 * log10(energy) - 0.691;
                            lufs = (double)(float)l;  // dsp/loudness.rs:123-125 keeps an f32
                        } else {
                            v1 = c1;
                            v2 = c2;
                            v3 = c3;
                            v4 = c4;
                        }
                        update(activity, reliability, blen);
                        mk_lin = db_to_lin(mk);
                        if (rows) {
                            rows[((size_t)0 * n_rows + clk.blk) * stride] = (float)mk;
                            rows[((size_t)1 * n_rows + clk.blk) * stride] = (float)score;
                            rows[((size_t)2 * n_rows + clk.blk) * stride] = (float)rel;
                        }
                        sq_in = 0.0;
                        full = tail = 0.0;
                        off = 0;
                        sub = 0;
                        clk.advance();
                    }
                }
            }
            store_tile(x + (size_t)t0 * stride, stride, valid, y);
        };
        pipelined_tiles(len, issue, body);
        // chunks are whole blocks: the filter state that carries over is the committed one (already in c1..c4)
    }
};

// ---- lookahead limiter (dsp/limiter.rs:246-284) -----------------------------------------------------------------
// M: target gain of samples n..n+7 from the sliding maximum of |x| over [n-L, n] (exact: max is order
// independent).  The 8 windows share [n+7-L, n]; each adds a few samples on the left and on the right.
constexpr int kLimGroup = 32;  // outputs per limiter-map thread: L + 32 loads for 32 windows

// x(m) = |input| of sample n_first + m, m in [-L, kLimGroup): FAST = the whole span is inside the ring slot
// history (no wrap, no samples before the render's start, no ragged tail), so it is one pointer walk.
template <bool FAST>
struct LimiterSpan {
    const float* ring;   // this stream's column of the ring
    const float* first;  // FAST: &x(-L)
    size_t stride;
    int ring_rows, row0, t0, n_first, valid, L;
    mutable bool nan_seen = false;  // a NaN was loaded: the windows that hold it take nan_aware_window (afsim_stages.h)
    AF_HD float operator()(int m) const {
        float v;
        if (FAST)
            v = fabsf(first[(size_t)(m + L) * stride]);
        else if (n_first + m < 0 || m >= valid)
            v = 0.0f;
        else
            v = fabsf(ring[(size_t)ring_row(row0, t0 + m, ring_rows) * stride]);
        nan_seen = nan_seen || v != v;
        return v;
    }
};

template <bool FAST>
AF_HD void limiter_windows(const LimiterSpan<FAST>& x, int L, float (&win)[kLimGroup]) {
    constexpr int G = kLimGroup;
    if (L >= G) {
        // window j = [j-L, j] = (suffix of the samples before the group, from j-L) U (prefix of the group, to j)
        float run = 0.0f;
        for (int m = -1; m >= G - L; --m) run = fmaxf(run, x(m));
        float left[G];
#pragma unroll
        for (int j = G - 1; j >= 0; --j) {
            run = fmaxf(run, x(j - L));
            left[j] = run;
        }
        float prefix = 0.0f;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            prefix = fmaxf(prefix, x(j));
            win[j] = fmaxf(left[j], prefix);
        }
    } else {
#pragma unroll 4
        for (int j = 0; j < G; ++j) {
            float w = 0.0f;
            for (int m = j - L; m <= j; ++m) w = fmaxf(w, x(m));
            win[j] = w;
        }
    }
    if (x.nan_seen) {  // rare: redo the group with the reference queue's NaN semantics
        for (int j = 0; j < G; ++j) {
            auto xj = [&](int m) { return x(j + m); };
            win[j] = nan_aware_window(xj, L);
        }
    }
}

AF_HD void limiter_targets(const float* in_ring, size_t stride, int ring_rows, int row0, int t0, int n_first, int valid,
                           int L, double ceil_lin, double* target_out /* column at t0 */) {
    constexpr int G = kLimGroup;
    float win[G];
    if (valid == G && n_first - L >= 0 && row0 + t0 - L >= 0) {
        LimiterSpan<true> x{in_ring, in_ring + (size_t)(row0 + t0 - L) * stride, stride, ring_rows, row0, t0, n_first, valid, L};
        limiter_windows(x, L, win);
    } else {
        LimiterSpan<false> x{in_ring, in_ring, stride, ring_rows, row0, t0, n_first, valid, L};
        limiter_windows(x, L, win);
    }
#pragma unroll
    for (int j = 0; j < G; ++j) {
        if (j < valid) {
            const double peak = (double)win[j];
            target_out[(size_t)j * stride] = peak > ceil_lin ? ceil_lin / peak : 1.0;
        }
    }
}

struct LimiterR {
    double g, min_g;
    AF_HD void init() {
        g = 1.0;
        min_g = 1.0;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
        io.f64(g);
        io.f64(min_g);
    }
    // target (w0 column at chunk start), delayed input from the ring, output column at chunk start
    AF_HD void run(const double* target, const float* in_ring, float* out, size_t stride, int ring_rows, int row0, int n0,
                   int len, int L, double ceil_lin, double rel, Staging stg) {
        constexpr int U = kGroup;
        const double one_m_rel = 1.0 - rel;
        const StageRing<double> stgt = stg.ring<double>();
        const StageRing<float> sdel = stg.ring<float>();
        auto issue = [&](int k, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = k * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || t0 + u < len) {
                    async_copy(stgt.at(k, u), target + (size_t)(t0 + u) * stride);
                    if (n0 + t0 + u >= L)
                        async_copy(sdel.at(k, u), in_ring + (size_t)ring_row(row0, t0 + u - L, ring_rows) * stride);
                    else
                        *sdel.at(k, u) = 0.0f;
                }
            }
        };
        auto body = [&](int k, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = k * U;
            const int valid = FULL ? U : len - t0;
            double tgt[U];
            float delayed[U], y[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                tgt[u] = (FULL || u < valid) ? *stgt.at(k, u) : 1.0;
                delayed[u] = (FULL || u < valid) ? *sdel.at(k, u) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || u < valid) {
                    if (tgt[u] < g)
                        g = tgt[u];
                    else
                        g = rel * g + one_m_rel * tgt[u];
                    min_g = fmin(min_g, g);
                }
                y[u] = (float)clampd((double)delayed[u] * g, -ceil_lin, ceil_lin);
            }
            store_tile(out + (size_t)t0 * stride, stride, valid, y);
        };
        pipelined_tiles(len, issue, body);
    }
    AF_HD float peak_reduction_db() const { return min_g < 1.0 ? (float)(-lin_to_db(min_g, 1e-10)) : 0.0f; }
};

// ---- 4x true-peak FIR as a map: peaks of samples n..n+7 of a ring column ------------------------------------------
template <typename FIR>
AF_HD void fir_group_peaks(const float* ring, size_t stride, int ring_rows, int row0, int t0, int n_first, int valid,
                           const FIR& fir, float (&peak)[kFirChunk]) {
    float win[kFirWin];
    if (valid == kFirChunk && n_first >= 31 && row0 + t0 >= 31) {  // whole 39-sample span inside the slot history
        const float* first = ring + (size_t)(row0 + t0 - 31) * stride;
#pragma unroll
        for (int i = 0; i < kFirWin; ++i) {
            const float v = first[(size_t)i * stride];
            win[i] = af_finite(v) ? v : 0.0f;  // dsp/true_peak.rs:211,343 sanitise
        }
    } else {
#pragma unroll
        for (int i = 0; i < kFirWin; ++i) {
            const int m = i - 31;  // sample n_first + m
            float v = 0.0f;
            if (n_first + m >= 0 && m < valid) v = ring[(size_t)ring_row(row0, t0 + m, ring_rows) * stride];
            win[i] = af_finite(v) ? v : 0.0f;
        }
    }
    fir8_peaks(win, fir, peak);
}

// ---- true-peak limiter gain recurrence + output statistics (dsp/true_peak.rs:337-378, python_api.rs:529-575) -------
struct TpR {
    float g, min_g;
    uint32_t events;
    bool limited;
    double sum_out, blk_out;
    float peak_out;
    bool non_finite;

    AF_HD void init() {
        g = 1.0f;
        min_g = 1.0f;
        events = 0;
        limited = false;
        sum_out = blk_out = 0.0;
        peak_out = 0.0f;
        non_finite = false;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
        io.f32(g);
        io.f32(min_g);
        io.u32(events);
        io.flag(limited);
        io.f64(sum_out);
        io.f64(blk_out);
        io.f32(peak_out);
        io.flag(non_finite);
    }
    // itp: target gains from the input true peaks (column at chunk start); in_ring: limiter output ring (the 20-sample delay
    // reads it); out: column at chunk start; audio: nullable
    AF_HD void run(const float* itp, const float* in_ring, float* out, float* audio, size_t stride, int ring_rows, int row0,
                   int n0, int len, float ceil_lin, float rel, BlockClock clk, float* rows_out, Staging stg) {
        constexpr int U = kGroup;
        const float one_m_rel = 1.0f - rel;
        const StageRing<float> spk = stg.ring<float>();
        const StageRing<float> sdel = stg.ring<float>();
        auto issue = [&](int k, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = k * U;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (FULL || t0 + u < len) {
                    async_copy(spk.at(k, u), itp + (size_t)(t0 + u) * stride);
                    if (n0 + t0 + u >= kTpDelay)
                        async_copy(sdel.at(k, u), in_ring + (size_t)ring_row(row0, t0 + u - kTpDelay, ring_rows) * stride);
                    else
                        *sdel.at(k, u) = 0.0f;
                }
            }
        };
        auto body = [&](int k, auto full) {
            constexpr bool FULL = decltype(full)::value;
            const int t0 = k * U;
            const int valid = FULL ? U : len - t0;
            float pk[U], delayed[U], y[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                pk[u] = (FULL || u < valid) ? *spk.at(k, u) : 0.0f;
                const float v = (FULL || u < valid) ? *sdel.at(k, u) : 0.0f;
                delayed[u] = af_finite(v) ? v : 0.0f;
            }
            auto walk = [&](auto may_end) {  // may_end: an analysis block can end inside this tile
                constexpr bool CHECK = decltype(may_end)::value;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    y[u] = 0.0f;
                    if (FULL || u < valid) {
                        const float target = pk[u];
                        if (target < g) {
                            g = target;
                            limited = true;
                        } else {
                            g = rel * g + one_m_rel * target;
                        }
                        min_g = fminf(min_g, g);
                        float o = clampf(delayed[u] * g, -ceil_lin, ceil_lin);
                        if (!af_finite(o)) o = 0.0f;
                        y[u] = o;
                        // output statistics (python_api.rs:529-575)
                        peak_out = fmaxf(peak_out, fabsf(o));
                        const double sq = (double)o * (double)o;
                        sum_out += sq;
                        blk_out += sq;
                        if (CHECK && clk.at_end(n0 + t0 + u)) {
                            events += limited ? 1u : 0u;
                            limited = false;
                            const float rms = (float)sqrt(blk_out / (double)clk.block_len(n0 + t0 + u));
                            rows_out[(size_t)clk.blk * stride] = lin_to_db_f32(rms);
                            blk_out = 0.0;
                            clk.advance();
                        }
                    }
                }
            };
            if (FULL && clk.ends_after(n0 + t0, U))
                walk(TileRagged());
            else
                walk(TileFull());
            store_tile(out + (size_t)t0 * stride, stride, valid, y);
            if (audio) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (u < valid) audio[t0 + u] = y[u];
            }
        };
        pipelined_tiles(len, issue, body);
    }
    AF_HD float peak_reduction_db() const { return min_g >= 1.0f ? 0.0f : -20.0f * af_log10_f32(fmaxf(min_g, 1e-10f)); }
};

}  // namespace afsim
