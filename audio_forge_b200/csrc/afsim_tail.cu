// afsim_tail.cu -- the chain's tail as ONE SM-local kernel: sample limiter (dsp/limiter.rs:246-284) -> 4x true-peak
// limiter (dsp/true_peak.rs:337-378) -> true-peak detector (dsp/true_peak.rs:208-218) + output statistics
// (audio/processor/python_api.rs:529-575).
//
// The R/M split set runs this as five kernels (k_lim_m, k_lim_r, k_tp_fir_in, k_tp_r, k_tp_fir_out) that hand
// every intermediate through HBM rings: ~60 B of DRAM traffic per stream-sample for 4 B of input, and two
// one-warp-per-SM serial kernels whose launch time sets the pipeline period.  Here one CTA owns 32 streams (one lane
// column of the stream-minor rings) for a whole chunk and keeps everything between the input and the statistics in
// shared memory:
//
//   producer warp   TMA (cp.async.bulk.tensor.2d, tile = 32 time rows x 32 streams of the f32 input ring, mbarrier
//                   complete_tx) into a time-circular x ring that also holds the L-sample lookback of the limiter
//   map warps (4)   per 32-row sub-tile, 8 rows each:  LIM-M  sliding-window maximum + f64 target gain
//                                                      FIR-IN 4x polyphase FIR over the limiter output -> input true
//                                                             peaks -> f32 target gains of the true-peak limiter
//                                                      FIR-OUT detector FIR over the true-peak limiter's output
//   serial warps    LIM-R (lane = stream: g <- t < g ? t : r g + (1 - r) t, delayed sample x gain, clamp) and
//                   TP-R (the same recurrence in f32, 20-sample delay, output statistics, per-block rows)
//
// Stages are linked by mbarriers in shared memory (one phase per finished sub-tile; waiting warps are suspended by the
// hardware), the x ring by the TMA's complete_tx mbarriers.  Per sample every operation and its order are
// those of the split kernels (afsim_split.h) -- the results are bit-identical to them and to the fused stages
// (tests/test_gpu_tail.py) -- only the hand-off medium changes.  Nothing but the statistics (and the audio, when the
// caller asked for it) leaves the SM: DRAM traffic is the 4 B read of the input per stream-sample.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "afsim_kernels.h"
#include "afsim_render.h"
#include "afsim_tail.h"

namespace afsim {

namespace {

__constant__ float c_tail_fir[4][32] = {
#include "true_peak_fir.inc"
};

constexpr int kSub = 32;  // time rows per sub-tile (= TMA box rows)
constexpr int kTG = 3;    // target-gain slots (f64, 8 KB each)
constexpr int kCO = 4;    // limiter-output ring, sub-tiles (one of them is history for the FIR / the 20-sample delay)
constexpr int kTT = 3;    // true-peak target slots
constexpr int kCY = 4;    // true-peak limiter output ring, sub-tiles
constexpr int kMaxCX = 32;
constexpr uint32_t kSpinLimit = 1u << 21;  // failed try_waits (~90 ns each: a fifth of a second) before the watchdog fires

struct TailCtl {  // shared-memory control block
    unsigned long long full[kMaxCX];  // x sub-tile landed (TMA complete_tx)
    // stage hand-offs: barrier [j & 3] completes its (j >> 2)-th phase when the stage has finished sub-tile j.  Waiting is
    // mbarrier.try_wait -- the warp is suspended by the hardware, not polling: with flag polling the polls of the waiting
    // warps were 45 % of all instructions this kernel issued (ncu, r02 passes D - F).  Depth 4 is enough: no stage can
    // finish sub-tile j + 4 before every waiter for its sub-tile j has passed (each is gated by a ring of <= 4 sub-tiles).
    unsigned long long lim_m[4];    // 2 arrivals (LIM-M warps 2, 7) -> LIM-R
    unsigned long long lim_r[4];    // 1 arrival  (LIM-R warp)      -> FIR-IN warps, LIM-M (target slot free)
    unsigned long long fir_in[4];   // 4 arrivals (warps 3 .. 6)    -> TP-R, LIM-R (limiter-output rows free)
    unsigned long long tp_r[4];     // 1 arrival  (TP-R warp)       -> FIR-OUT warps, FIR-IN (target slot free), LIM-R
    unsigned long long fir_out[4];  // 4 arrivals (warps 3 .. 6)    -> TP-R (output rows free)
    int error;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Waits for the phase with the given parity.  The loop is a handful of instructions per failed try (the hardware parks
// the warp for a short time slice per try_wait, ~90 ns measured, whatever the hint says -- so tries are frequent, and a
// 18-instruction C++ loop around them was 41 % of all instructions issued).  A stuck pipeline (a bug) must never hang the
// GPU: after kSpinLimit failed tries the CTA's sticky error flag is set and every later wait gives up at once.
__device__ __noinline__ void mbar_wait(unsigned long long* bar, uint32_t parity, TailCtl* ctl) {
    uint32_t tries;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
        "mov.u32 n, 0;\n"
        "TAIL_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "@p bra TAIL_WAIT_DONE;\n\t"
        "add.u32 n, n, 1;\n\t"
        "setp.lt.u32 p, n, %4;\n\t"
        "@p bra TAIL_WAIT_LOOP;\n"
        "TAIL_WAIT_DONE:\n\t"
        "mov.u32 %0, n;\n\t}"
        : "=r"(tries)
        : "r"(smem_u32(bar)), "r"(parity), "r"(2000u), "r"(ctl->error ? 1u : kSpinLimit)
        : "memory");
    if (tries >= kSpinLimit) st_release(&ctl->error, 1);
}
// "stage finished sub-tile j": wait for / signal phase j >> 2 of barrier j & 3
__device__ __forceinline__ void wait_done(unsigned long long* bars, int j, TailCtl* ctl) {
    mbar_wait(&bars[j & 3], (uint32_t)((j >> 2) & 1), ctl);
}
__device__ __forceinline__ void signal_done(unsigned long long* bars, int j) { mbar_arrive(&bars[j & 3]); }
// 2-D tile of the [ring_rows][S_pad] f32 ring: c0 = first stream (column), c1 = first ring row
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// per-block output level row (python_api.rs:54-56,570-575): one out-of-line instance of the sqrt / log10 code
__device__ __noinline__ void store_output_row(float* dst, double block_sum, int block_len) {
    const float rms = (float)sqrt(block_sum / (double)block_len);
    *dst = lin_to_db_f32(rms);
}

struct TailSmem {
    TailCtl* ctl;
    float* xs;    // [cx * 32][32]   x ring (TMA destination)
    double* tg;   // [kTG * 32][32]  limiter target gains
    float* ol;    // [kCO * 32][32]  limiter output
    float* tt;    // [kTT * 32][32]  true-peak limiter target gains
    float* ys;    // [kCY * 32][32]  true-peak limiter output
};

__host__ __device__ inline size_t tail_smem_bytes(int cx) {
    return 1024 + (size_t)cx * kSub * 32 * 4 + (size_t)kTG * kSub * 32 * 8 + (size_t)(kCO + kTT + kCY) * kSub * 32 * 4;
}

extern __shared__ __align__(1024) unsigned char tail_smem_raw[];

__device__ __forceinline__ TailSmem carve(int cx) {
    TailSmem sm;
    unsigned char* p = tail_smem_raw;
    sm.ctl = reinterpret_cast<TailCtl*>(p);
    p += 1024;
    sm.xs = reinterpret_cast<float*>(p);
    p += (size_t)cx * kSub * 32 * 4;
    sm.tg = reinterpret_cast<double*>(p);
    p += (size_t)kTG * kSub * 32 * 8;
    sm.ol = reinterpret_cast<float*>(p);
    p += (size_t)kCO * kSub * 32 * 4;
    sm.tt = reinterpret_cast<float*>(p);
    p += (size_t)kTT * kSub * 32 * 4;
    sm.ys = reinterpret_cast<float*>(p);
    return sm;
}

// ---- map tasks (lane = stream) ------------------------------------------------------------------------------------------

// LIM-M, G rows per task: windows [r - L, r] of |x| for the rows r = base .. base + G - 1 (dsp/limiter.rs:253-262; max is
// order independent, so the part the G windows share is scanned once: L + G loads for G outputs), then the target gain
// in f64.  The scan runs on NaN-PROPAGATING maxima (max.NaN.f32): without a NaN in the span they equal
// fmaxf bit for bit, and a NaN result sends the sub-tile to the reference queue's NaN semantics (nan_aware_window).
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
template <int G>
__device__ __forceinline__ void task_lim_m(const float* xs, int xmask, int lane, int base, int L, double ceil_lin,
                                           double* tg /* the slot row of the task's first row */) {
    auto x = [&](int m) { return fabsf(xs[(size_t)((base + m) & xmask) * 32 + lane]); };
    auto store_target = [&](int j, float w) {
        const double peak = (double)w;
        tg[(size_t)j * 32 + lane] = peak > ceil_lin ? ceil_lin / peak : 1.0;
    };
    // The suffix parts of the 32 windows are parked as floats in the slot itself (each thread its own column; the double
    // written later covers the float): small rolled loops instead of a 32-register array -- the kernel's seven warps
    // run four different code paths at once, and code size is what the instruction cache feels.
    float* scratch = reinterpret_cast<float*>(tg);
    bool nan_seen = false;
    if (L >= G) {
        float run = 0.0f;
#pragma unroll 4
        for (int m = -1; m >= G - L; --m) run = max_nan(run, x(m));
#pragma unroll 4
        for (int j = G - 1; j >= 0; --j) {  // suffix part of window j: the samples before the sub-tile, from j - L on
            run = max_nan(run, x(j - L));
            scratch[(size_t)j * 64 + 2 * lane] = run;
        }
        float prefix = 0.0f;
#pragma unroll 4
        for (int j = 0; j < G; ++j) {       // prefix part: the sub-tile's own samples up to j
            prefix = max_nan(prefix, x(j));
            const float w = max_nan(scratch[(size_t)j * 64 + 2 * lane], prefix);
            nan_seen = nan_seen || w != w;
            scratch[(size_t)j * 64 + 2 * lane] = w;
        }
    } else {
#pragma unroll 1
        for (int j = 0; j < G; ++j) {
            float w = 0.0f;
            for (int m = j - L; m <= j; ++m) w = max_nan(w, x(m));
            nan_seen = nan_seen || w != w;
            scratch[(size_t)j * 64 + 2 * lane] = w;
        }
    }
    // the f64 divisions, eight independent ones in flight: one division is a ~15-deep dependent FP64 chain, and this warp
    // is alone on its role -- rolled two at a time it was the slowest stage of the whole pipeline (84 % busy, ncu pass H)
#pragma unroll 1
    for (int j0 = 0; j0 < G; j0 += 8) {
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = scratch[(size_t)(j0 + i) * 64 + 2 * lane];
#pragma unroll
        for (int i = 0; i < 8; ++i) store_target(j0 + i, w[i]);
    }
    if (nan_seen) {  // rare: a NaN inside a lookback -- every row again with the reference queue's semantics
#pragma unroll 1
        for (int j = 0; j < G; ++j) {
            auto xj = [&](int m) { return x(j + m); };
            store_target(j, nan_aware_window(xj, L));
        }
    }
}

// 4x polyphase FIR peaks of rows base .. base + 7 of a ring (taps in the reference's order, dsp/true_peak.rs:173-186).
// The rings hold SANITISED samples (non-finite -> 0 at the write, dsp/true_peak.rs:211,343), so the window loads are plain.
__device__ __forceinline__ void task_fir(const float* ring, int mask, int lane, int base, float (&pk)[kFirChunk]) {
    float win[kFirWin];
#pragma unroll
    for (int i = 0; i < kFirWin; ++i) win[i] = ring[(size_t)((base - 31 + i) & mask) * 32 + lane];
    fir8_peaks(win, c_tail_fir, pk);
}

}  // namespace

// One CTA = 32 streams x one chunk, 8 warps:
//   warp 0  LIM-R + TMA producer      warp 1  TP-R + output statistics      warps 2, 7  LIM-M (16 rows of a sub-tile each)
//   warps 3..6  FIR units of 8 rows: one FIR-IN and one FIR-OUT unit per warp and step
__global__ void __launch_bounds__(kTailThreads, 2)
k_tail(BatchArgs a, ChunkArgs ck, const __grid_constant__ CUtensorMap x_map, int cx, int* err_out) {
    if (err_out && *reinterpret_cast<volatile int*>(err_out)) return;  // an earlier launch's watchdog fired: do not pile up waits
    const TailSmem sm = carve(cx);
    TailCtl* ctl = sm.ctl;
    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
    const int s_base = (int)blockIdx.x * 32, s = s_base + lane;
    const size_t stride = (size_t)a.stride;
    const int L = a.lookahead, Lt = (L + kSub - 1) / kSub;
    const int n_sub = (ck.len + kSub - 1) / kSub;
    const int xmask = cx * kSub - 1, omask = kCO * kSub - 1, ymask = kCY * kSub - 1;
    // x ring: sample r of the chunk (r >= -Lt * 32: the lookback) lives at row (r + xoff) & xmask, i.e. sub-tile j in slot
    // (j + Lt) mod cx -- its (j + Lt) / cx-th use, which is the parity its mbarrier is waited on with
    const int xoff = Lt * kSub;
    auto x_slot = [&](int j) { return (j + Lt) & (cx - 1); };
    auto x_parity = [&](int j) { return (uint32_t)(((j + Lt) / cx) & 1); };
    const bool first_chunk = ck.n0 == 0;

    // ---- init: barriers, counters, histories -----------------------------------------------------------------------
    if (threadIdx.x == 0) {
        for (int i = 0; i < cx; ++i) mbar_init(&ctl->full[i], 1);
        for (int i = 0; i < 4; ++i) {
            mbar_init(&ctl->lim_m[i], 2);
            mbar_init(&ctl->lim_r[i], 1);
            mbar_init(&ctl->fir_in[i], 4);
            mbar_init(&ctl->tp_r[i], 1);
            mbar_init(&ctl->fir_out[i], 4);
        }
        ctl->error = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // sub-tile -1 of the limiter-output / true-peak-output rings: the previous chunk's last 32 rows (parked), or silence
    for (int i = (int)threadIdx.x; i < kSub * 32; i += kTailThreads) {
        const int r = i >> 5, c = i & 31;
        float ho = 0.0f, hy = 0.0f;
        if (!first_chunk) {
            ho = a.tail_hist[(size_t)r * stride + s_base + c];
            hy = a.tail_hist[(size_t)(kSub + r) * stride + s_base + c];
        }
        sm.ol[(size_t)((r - kSub) & omask) * 32 + c] = ho;
        sm.ys[(size_t)((r - kSub) & ymask) * 32 + c] = hy;
    }
    if (first_chunk)  // samples before the render's start are silence (dsp/limiter.rs:117-131: the delay line starts at zero)
        for (int i = (int)threadIdx.x; i < xoff * 32; i += kTailThreads) sm.xs[i] = 0.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes above vs the TMA writes that reuse the slots
    __syncthreads();

    if (warp == 0) {
        // ---- LIM-R: gain recurrence of the sample limiter, lane = stream; lane 0 also feeds the x ring -----------------------
        auto load_subtile = [&](int j) {  // one thread: sub-tile j -> its slot (the slot's previous tenant is no longer needed)
            const int slot = x_slot(j);
            if (j < 0 && first_chunk) {
                mbar_arrive(&ctl->full[slot]);  // zero-filled above; keeps the phase sequence uniform
            } else {
                int row = ck.row0 + j * kSub;
                if (row < 0) row += a.ring_rows;
                mbar_expect_tx(&ctl->full[slot], kSub * 32 * 4);
                tma_load_2d(sm.xs + (size_t)slot * kSub * 32, &x_map, s_base, row, &ctl->full[slot]);
            }
        };
        // slot (j mod cx) is free once LIM-R finished sub-tile j - cx + Lt: the first cx sub-tiles need no wait
        if (lane == 0)
            for (int j = -Lt; j < n_sub && j < cx - Lt; ++j) load_subtile(j);
        const CandidateParams& p = stream_params(a, s);
        const double ceil_lin = p.l_ceil, rel = p.l_release, one_m_rel = 1.0 - rel;
        LimiterR st;
        if (first_chunk) {
            st.init();
        } else {
            StateIO<false> io{a.st_lim + s, stride};
            st.sync(io);
        }
        double g = st.g, min_g = st.min_g;
        for (int j = 0; j < n_sub; ++j) {
            const int valid = ck.len - j * kSub < kSub ? ck.len - j * kSub : kSub;
            wait_done(ctl->lim_m, j, ctl);
            const int k = j - kCO + 1;  // the ring rows about to be overwritten were last read by FIR-IN(k) and TP-R(k)
            if (k >= 0) {
                wait_done(ctl->fir_in, k, ctl);
                wait_done(ctl->tp_r, k, ctl);
            }
            const double* tg = sm.tg + (size_t)(j % kTG) * kSub * 32 + lane;
#pragma unroll 1
            for (int u0 = 0; u0 < kSub; u0 += 8) {
                double tgt[8];
                float delayed[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    tgt[u] = tg[(size_t)(u0 + u) * 32];
                    delayed[u] = sm.xs[(size_t)((j * kSub + u0 + u - L + xoff) & xmask) * 32 + lane];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (u0 + u < valid) {
                        if (tgt[u] < g)
                            g = tgt[u];
                        else
                            g = rel * g + one_m_rel * tgt[u];
                        min_g = fmin(min_g, g);
                        const float y = (float)clampd((double)delayed[u] * g, -ceil_lin, ceil_lin);
                        // both readers sanitise what they read (dsp/true_peak.rs:211,343): done once, here
                        sm.ol[(size_t)((j * kSub + u0 + u) & omask) * 32 + lane] = af_finite(y) ? y : 0.0f;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                signal_done(ctl->lim_r, j);
                const int jn = j + cx - Lt;  // its slot held sub-tile j - Lt: nobody needs that any more
                if (jn < n_sub) load_subtile(jn);
            }
        }
        st.g = g;
        st.min_g = min_g;
        if (ck.n0 + ck.len >= a.n_samples) {
            a.accum[s].limiter_gr_db = st.peak_reduction_db();
        } else {
            StateIO<true> io{a.st_lim + s, stride};
            st.sync(io);
        }
    } else if (warp == 1) {
        // ---- TP-R: gain recurrence of the true-peak limiter + output statistics, lane = stream -------------------------------
        const CandidateParams& p = stream_params(a, s);
        const float ceil_lin = p.tp_ceil, rel = p.tp_release, one_m_rel = 1.0f - rel;
        TpR st;
        if (first_chunk) {
            st.init();
        } else {
            StateIO<false> io{a.st_tp + s, stride};
            st.sync(io);
        }
        BlockClock clk;
        clk.init(a.block_samples, a.n_samples, ck.n0);
        float* rows_out = a.rows + (size_t)1 * a.n_rows * stride + s;
        for (int j = 0; j < n_sub; ++j) {
            const int valid = ck.len - j * kSub < kSub ? ck.len - j * kSub : kSub;
            wait_done(ctl->fir_in, j, ctl);
            const int k = j - kCY + 1;  // last reader of the output ring rows about to be overwritten: FIR-OUT(k)
            if (k >= 0) wait_done(ctl->fir_out, k, ctl);
            const float* tt = sm.tt + (size_t)(j % kTT) * kSub * 32 + lane;
#pragma unroll 1
            for (int u0 = 0; u0 < kSub; u0 += 8) {
                float pk[8], delayed[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    pk[u] = tt[(size_t)(u0 + u) * 32];
                    delayed[u] = sm.ol[(size_t)((j * kSub + u0 + u - kTpDelay) & omask) * 32 + lane];  // sanitised at the write
                }
                const int n_first = ck.n0 + j * kSub + u0;
                auto walk = [&](auto may_end) {
                    constexpr bool CHECK = decltype(may_end)::value;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (u0 + u < valid) {
                            const float target = pk[u];
                            if (target < st.g) {
                                st.g = target;
                                st.limited = true;
                            } else {
                                st.g = rel * st.g + one_m_rel * target;
                            }
                            st.min_g = fminf(st.min_g, st.g);
                            float o = clampf(delayed[u] * st.g, -ceil_lin, ceil_lin);
                            if (!af_finite(o)) o = 0.0f;
                            sm.ys[(size_t)((j * kSub + u0 + u) & ymask) * 32 + lane] = o;
                            st.peak_out = fmaxf(st.peak_out, fabsf(o));
                            const double sq = (double)o * (double)o;
                            st.sum_out += sq;
                            st.blk_out += sq;
                            if (CHECK && clk.at_end(n_first + u)) {
                                st.events += st.limited ? 1u : 0u;
                                st.limited = false;
                                store_output_row(rows_out + (size_t)clk.blk * stride, st.blk_out, clk.block_len(n_first + u));
                                st.blk_out = 0.0;
                                clk.advance();
                            }
                        }
                    }
                };
                if (clk.ends_after(n_first, 8))  // no analysis block (nor the signal) ends inside these 8 samples
                    walk(TileRagged());
                else
                    walk(TileFull());
            }
            __syncwarp();
            if (lane == 0) signal_done(ctl->tp_r, j);
        }
        if (ck.n0 + ck.len >= a.n_samples) {
            StreamAccum& acc = a.accum[s];
            acc.sum_out = st.sum_out;
            acc.peak_out = st.peak_out;
            acc.non_finite = st.non_finite ? 1u : 0u;
            acc.tp_gr_db = st.peak_reduction_db();
            acc.events = st.events;
        } else {
            StateIO<true> io{a.st_tp + s, stride};
            st.sync(io);
        }
    } else if (warp == 2 || warp == 7) {
        // ---- LIM-M: two warps, 16 rows of every sub-tile each (alone this role was the slowest stage: 80 % busy) -----------
        const int half = warp == 2 ? 0 : 1;
        const double l_ceil = a.map_tab[(size_t)MT_L_CEIL * stride + s];
        for (int j = 0; j < n_sub; ++j) {
            if (j == 0)
                for (int h = -Lt; h < 0; ++h) mbar_wait(&ctl->full[x_slot(h)], x_parity(h), ctl);
            mbar_wait(&ctl->full[x_slot(j)], x_parity(j), ctl);
            if (j - kTG >= 0) wait_done(ctl->lim_r, j - kTG, ctl);  // the slot's previous targets were consumed
            task_lim_m<kSub / 2>(sm.xs, xmask, lane, j * kSub + 16 * half + xoff, L, l_ceil,
                                 sm.tg + (size_t)((j % kTG) * kSub + 16 * half) * 32);
            __syncwarp();
            if (lane == 0) signal_done(ctl->lim_m, j);
        }
    } else {
        // ---- FIR units: unit q = rows [8q, 8q + 8) of a sub-tile.  Per step t: the FIR-IN units of sub-tile t - 1 (they feed
        // the serial TP-R warp: first), then the FIR-OUT unit of sub-tile t - 3 -- TWO steps behind the FIR-IN of the same
        // sub-tile, so that TP-R has a whole step to turn its targets into output rows and no FIR warp ever waits for it.
        // Warp 3 + m owns unit m of both kinds (the serial warps issue so little that the schedulers hosting them need no
        // relief: a 3 / 2 / 2 / 1 split only made warp 3 the straggler).  ONE instance of the unrolled FIR serves both kinds.
        const int m = warp - 3;  // warps 3 .. 6
        const int in_first = m, in_count = 1;  // one FIR-IN and one FIR-OUT unit per warp and step
        const float tp_ceil = (float)a.map_tab[(size_t)MT_TP_CEIL * stride + s];
        float* audio = (a.audio && s < a.n_streams) ? a.audio + a.audio_off[s] + ck.n0 : nullptr;
        float max_in = 0.0f, max_out = 0.0f;
        for (int t = 1; t < n_sub + 3; ++t) {
#pragma unroll 1
            for (int task = 0; task <= in_count; ++task) {  // tasks 0 .. in_count - 1: the IN units; the last one: the OUT unit
                const bool is_out = task == in_count;
                const int j = is_out ? t - 3 : t - 1;
                if (j < 0 || j >= n_sub) continue;
                const int q = is_out ? m : in_first + task;
                const int base = j * kSub + 8 * q;
                const int valid = ck.len - base;
                if (is_out) {
                    wait_done(ctl->tp_r, j, ctl);
                } else if (task == 0) {
                    wait_done(ctl->lim_r, j, ctl);
                    if (j - kTT >= 0) wait_done(ctl->tp_r, j - kTT, ctl);  // the slot's previous targets were consumed
                }
                float pk[kFirChunk];
                task_fir(is_out ? sm.ys : sm.ol, is_out ? ymask : omask, lane, base, pk);
                if (is_out) {
#pragma unroll
                    for (int i = 0; i < kFirChunk; ++i)
                        if (i < valid) max_out = fmaxf(max_out, pk[i]);
                    if (audio) {
#pragma unroll
                        for (int i = 0; i < kFirChunk; ++i)
                            if (i < valid) audio[base + i] = sm.ys[(size_t)((base + i) & ymask) * 32 + lane];
                    }
                    __syncwarp();
                    if (lane == 0) signal_done(ctl->fir_out, j);
                } else {
                    float* tt = sm.tt + (size_t)((j % kTT) * kSub + 8 * q) * 32 + lane;
#pragma unroll
                    for (int i = 0; i < kFirChunk; ++i) {
                        // feed-forward part of dsp/true_peak.rs:349-354
                        tt[(size_t)i * 32] = pk[i] > tp_ceil ? clampf((tp_ceil * 0.999f) / pk[i], 0.0f, 1.0f) : 1.0f;
                        if (i < valid) max_in = fmaxf(max_in, pk[i]);
                    }
                    __syncwarp();
                    if (lane == 0 && task == in_count - 1) signal_done(ctl->fir_in, j);  // one arrival per warp, after its last IN unit
                }
            }
        }
        // running maxima of the two oversamplers: order independent (python_api.rs:552-560) -> atomics
        if (in_count > 0) atomic_max_nonneg(&a.accum[s].peak_pre_tp, max_in);
        atomic_max_nonneg(&a.accum[s].peak_out_tp, max_out);
    }
    __syncthreads();
    // park the last 32 rows of both rings for the next chunk's FIR history and 20-sample delay
    if (ck.n0 + ck.len < a.n_samples)
        for (int i = (int)threadIdx.x; i < kSub * 32; i += kTailThreads) {
            const int r = i >> 5, c = i & 31;
            const int row = ck.len - kSub + r;
            a.tail_hist[(size_t)r * stride + s_base + c] = sm.ol[(size_t)(row & omask) * 32 + c];
            a.tail_hist[(size_t)(kSub + r) * stride + s_base + c] = sm.ys[(size_t)(row & ymask) * 32 + c];
        }
    if (threadIdx.x == 0 && ctl->error && err_out) atomicExch(err_out, 1);
}

// ---- host side ---------------------------------------------------------------------------------------------------------

int tail_ring_subtiles(int lookahead) {
    const int lt = (lookahead + kSub - 1) / kSub;
    int cx = 8;
    while (cx < lt + 4) cx <<= 1;
    return cx;
}

bool tail_supported(int lookahead) { return lookahead >= 1 && tail_ring_subtiles(lookahead) <= kMaxCX; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

cudaError_t tail_make_map(const float* ring, int ring_rows, int stride, TailMap* out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        const cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (err != cudaSuccess) return err;
        if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    static_assert(sizeof(TailMap) == sizeof(CUtensorMap), "TailMap must hold a CUtensorMap");
    const cuuint64_t dims[2] = {(cuuint64_t)stride, (cuuint64_t)ring_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)stride * sizeof(float)};
    const cuuint32_t box[2] = {32, (cuuint32_t)kSub};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult rc = encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ring),
                               dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t tail_configure() {
    return cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem_bytes(kMaxCX));
}

cudaError_t launch_tail(const BatchArgs& a, const ChunkArgs& ck, const TailMap& map, int* err_flag, cudaStream_t st) {
    const int cx = tail_ring_subtiles(a.lookahead);
    CUtensorMap m;
    memcpy(&m, &map, sizeof m);
    k_tail<<<dim3((unsigned)(a.stride / 32)), kTailThreads, tail_smem_bytes(cx), st>>>(a, ck, m, cx, err_flag);
    return cudaGetLastError();
}

}  // namespace afsim
