// afsim_kernels.cu -- sm_100a stage kernels of the batched chain simulator.
//
// One thread = one stream (candidate x passage pair); 32 consecutive streams = one warp, so every
// access to the time-major work buffers and the stream-minor state tables is a coalesced 128 / 256
// byte transaction.  Each kernel advances every stream of a batch over one chunk of samples for
// ONE stage of the chain (afsim_render.h holds the per-stream bodies, afsim_stages.h the
// arithmetic).  No tensor cores: nothing on this path is a dense contraction; the
// recurrences are FP64-issue bound, the true-peak FIR is FP32-FMA-issue bound.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true
//        -ftz=false (the reference never contracts a*b+c and runs with denormals on; the FIR uses
//        explicit __fmaf_rn).
#include <cuda_runtime.h>

#include <cstdlib>

#include "afsim_kernels.h"
#include "afsim_render.h"

namespace afsim {

__constant__ float c_fir[4][32] = {
#include "true_peak_fir.inc"
};

#define AF_STREAM_INDEX()                                         \
    const int s = (int)(blockIdx.x * blockDim.x + threadIdx.x);   \
    if (s >= a.n_streams) return

__global__ void __launch_bounds__(128) k_expand_deesser(BatchArgs a) {
    AF_STREAM_INDEX();
    body_expand_deesser(a, s);
}

__global__ void __launch_bounds__(128) k_input(BatchArgs a, ChunkArgs ck) {
    AF_STREAM_INDEX();
    body_input(a, ck, s);
}

// Input stage with the adaptive hum / rumble cleanup (afsim_cleanup.h): its own kernel, because its 26 phasor
// bins want ~250 registers and must not cost the plain input kernel its occupancy.
__global__ void __launch_bounds__(128) k_input_cleanup(BatchArgs a, ChunkArgs ck) {
    AF_STREAM_INDEX();
    body_input_cleanup(a, ck, s);
}

// The same stage with one WARP per stream (CleanupStageT<1>, afsim_cleanup.h): lane i owns hum bin i, everything
// else is computed redundantly by every lane and lane 0 stores it.  Used for the shared input stage of candidate
// sweeps (a handful of distinct passages), where this serial chain sets the wavefront's period.
__global__ void __launch_bounds__(32) k_input_cleanup_warp(BatchArgs a, ChunkArgs ck) {
    const int s = (int)blockIdx.x;
    if (s >= a.n_streams) return;
    const int lane = (int)threadIdx.x;
    const size_t stride = (size_t)a.stride;
    const CleanupConst& k = *a.cleanup;
    const bool gentle = a.input_stage == AF_INPUT_CLEANUP_GENTLE;
    InputStage st;
    CleanupStageT<1> cl;
    cl.init(k, lane < 2 * kHumBins ? lane : 2 * kHumBins - 1);
    if (ck.n0 == 0) {
        st.init();
    } else {
        StateIO<false> io{a.st_input + s, stride};
        st.sync(io);
        cl.sync(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    const Col out{a.buf_a + (size_t)ck.row0 * stride + s, stride};
    const float* src = a.signals + a.src_off[s] + ck.n0;
    float* rows_in = a.rows + s;
    // The block's samples go through shared memory, and the NEXT block's are already in flight (a coalesced load into
    // registers) while this one is walked: with a global load per sample in the two serial loops below, the warp
    // paid an exposed L2 / DRAM latency twice per sample -- 0.69 ms per 1440-sample chunk alone and 2.6x that inside
    // the wavefront, where the memory system is busy, which made this stage the wavefront's period.
    constexpr int kPerLane = (kInputBlock + 31) / 32;
    __shared__ float blk[kInputBlock];
    float next[kPerLane];
    auto prefetch = [&](int b0) {
        const int blen = ck.len - b0 < kInputBlock ? ck.len - b0 : kInputBlock;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {
            const int t = i * 32 + lane;
            next[i] = t < blen ? src[b0 + t] : 0.0f;
        }
    };
    prefetch(0);
    for (int b0 = 0; b0 < ck.len; b0 += kInputBlock) {
        const int blen = ck.len - b0 < kInputBlock ? ck.len - b0 : kInputBlock;
        __syncwarp();  // the previous block has been walked by every lane
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {
            const int t = i * 32 + lane;
            if (t < kInputBlock) blk[t] = af_finite(next[i]) ? next[i] : 0.0f;  // sanitised once
        }
        __syncwarp();
        if (b0 + kInputBlock < ck.len) prefetch(b0 + kInputBlock);
        for (int t = 0; t < blen; ++t) cl.analyze(blk[t], k, gentle);  // analyze_input on the raw (sanitised) block
        bool hum_detected;
        cl.begin_block(k, gentle, &hum_detected);
        for (int t = 0; t < blen; ++t) {
            const int n = ck.n0 + b0 + t;
            const float in = blk[t];
            const float dc = in - st.x1 + 0.995f * st.y1;  // routing.rs:832-836
            st.x1 = in;
            st.y1 = dc;
            float v = cl.process(dc, k);
            if (!af_finite(v)) v = 0.0f;
            const double sq = (double)v * (double)v;
            st.sum_in += sq;
            st.blk_in += sq;
            st.peak_in = fmaxf(st.peak_in, fabsf(v));
            if (lane == 0) out.set(b0 + t, v);
            if (clk.at_end(n)) {
                const float rms = (float)sqrt(st.blk_in / (double)clk.block_len(n));
                if (lane == 0) rows_in[(size_t)clk.blk * stride] = lin_to_db_f32(rms);
                st.blk_in = 0.0;
                clk.advance();
            }
        }
    }
    if (ck.n0 + ck.len >= a.n_samples) {
        if (lane == 0) {
            a.accum[s].sum_in = st.sum_in;
            a.accum[s].peak_in = st.peak_in;
        }
    } else {
        StateIO<true> io{a.st_input + s, stride};
        st.sync(io);
        cl.sync(io);
    }
}

template <int K>
__global__ void __launch_bounds__(128) k_eq(BatchArgs a, ChunkArgs ck, int first) {
    AF_STREAM_INDEX();
    body_eq<K>(a, ck, s, first);
}

__global__ void __launch_bounds__(128) k_compressor(BatchArgs a, ChunkArgs ck) {
    AF_STREAM_INDEX();
    body_compressor(a, ck, s);
}

__global__ void __launch_bounds__(128) k_compressor_shared(BatchArgs a, ChunkArgs ck) {
    AF_STREAM_INDEX();
    body_compressor_shared(a, ck, s);
}

__global__ void __launch_bounds__(128) k_limiter(BatchArgs a, ChunkArgs ck) {
    AF_STREAM_INDEX();
    body_limiter(a, ck, s);
}

template <bool LIMITER>
__global__ void __launch_bounds__(128) k_output(BatchArgs a, ChunkArgs ck) {
    AF_STREAM_INDEX();
    body_output<LIMITER>(a, ck, s, c_fir);
}

__global__ void __launch_bounds__(128) k_input_true_peak(BatchArgs a, ChunkArgs ck) {
    AF_STREAM_INDEX();
    body_input_true_peak(a, ck, s, c_fir);
}

// ---- split (R/M) path: serial recurrences, one thread per stream --------------------------------------------
// Each thread stages its inputs through shared memory with cp.async (Staging, afsim_split.h).
extern __shared__ __align__(16) unsigned char stage_smem[];
// Staged kernels run one warp per block (kRBlock lanes share the staging area): base and lane count are
// compile-time constants, so every staging address folds to `lane + constant`.  The de-esser's serial kernels also
// exist as direct variants (128-thread blocks reading global memory) for batches that fill the GPU by themselves.
#define AF_R_KERNEL(name, body)                                                         \
    __global__ void __launch_bounds__(kRBlock) name(BatchArgs a, ChunkArgs ck) {        \
        AF_STREAM_INDEX();                                                              \
        const Staging stg{stage_smem, kRBlock, (int)threadIdx.x, 0, 1};                 \
        body(a, ck, s, stg);                                                            \
    }
#define AF_R_KERNEL_DIRECT(name, body, minblocks)                                       \
    __global__ void __launch_bounds__(128, minblocks) name(BatchArgs a, ChunkArgs ck) { \
        AF_STREAM_INDEX();                                                              \
        const Staging stg{nullptr, 128, (int)threadIdx.x, 0, 0};                        \
        body(a, ck, s, stg);                                                            \
    }
AF_R_KERNEL(k_comp_r1, body_comp_r1)
AF_R_KERNEL(k_comp_r3, body_comp_r3)
AF_R_KERNEL(k_comp_r5, body_comp_r5)
AF_R_KERNEL(k_comp_r7, body_comp_r7)
AF_R_KERNEL(k_lim_r, body_lim_r)
AF_R_KERNEL(k_tp_r, body_tp_r)
AF_R_KERNEL(k_de_ra, body_de_ra)
AF_R_KERNEL_DIRECT(k_de_ra_direct, body_de_ra, 1)
// de-esser R_c1 (targets, hysteresis) and R_c3 (time-varying biquads): staged and direct variants
AF_R_KERNEL(k_de_rc, body_de_rc)
AF_R_KERNEL_DIRECT(k_de_rc_direct, body_de_rc, 4)
AF_R_KERNEL(k_de_rc1a, body_de_rc1a)
AF_R_KERNEL_DIRECT(k_de_rc1a_direct, body_de_rc1a, 4)
AF_R_KERNEL(k_de_rc1c, body_de_rc1c)
AF_R_KERNEL_DIRECT(k_de_rc1c_direct, body_de_rc1c, 4)
AF_R_KERNEL(k_de_rc3, body_de_rc3)
AF_R_KERNEL_DIRECT(k_de_rc3_direct, body_de_rc3, 4)

// ---- split path: maps, one thread per (stream, group of kGroup samples); blockIdx.y = group -------------------
// A block is kMapWarps warps over the SAME 32 streams and consecutive sample groups, so the overlapping
// rows that neighbouring groups read (31-sample FIR history, limiter window) hit in L1.
#define AF_M_KERNEL(name, GROUP, call)                                                         \
    __global__ void __launch_bounds__(32 * kMapWarps) name(BatchArgs a, ChunkArgs ck) {  \
        const int s = (int)(blockIdx.x * 32 + (threadIdx.x & 31));                       \
        if (s >= a.n_streams) return;                                                    \
        const int n_groups = (ck.len + (GROUP) - 1) / (GROUP);                           \
        for (int g = (int)(blockIdx.y * kMapWarps + (threadIdx.x >> 5)); g < n_groups;   \
             g += (int)gridDim.y * kMapWarps) {                                          \
            call;                                                                        \
        }                                                                                \
    }
AF_M_KERNEL(k_input_fanout, kFanoutGroup, body_input_fanout(a, ck, s, g))
AF_M_KERNEL(k_comp_m2, kCompMapGroup, body_comp_m2(a, ck, s, g))
AF_M_KERNEL(k_comp_m4, kCompMapGroup, body_comp_m4(a, ck, s, g))
AF_M_KERNEL(k_comp_m6, kCompMapGroup, body_comp_m6(a, ck, s, g))
AF_M_KERNEL(k_de_mb, kDeMapGroup, body_de_mb(a, ck, s, g))
AF_M_KERNEL(k_de_mc2, kDeRebuildGroup, body_de_mc2(a, ck, s, g))
AF_M_KERNEL(k_de_mc1b, kDeTargetGroup, body_de_mc1b(a, ck, s, g))
AF_M_KERNEL(k_lim_m, kLimGroup, body_lim_m(a, ck, s, g))
AF_M_KERNEL(k_tp_fir_in, kGroup, body_tp_fir_in(a, ck, s, g, c_fir))
AF_M_KERNEL(k_tp_fir_out, kGroup, body_tp_fir_out(a, ck, s, g, c_fir))

extern __shared__ __align__(16) float fin_smem[];

// One thread block per stream: sorts and percentiles of the per-block rows.
__global__ void __launch_bounds__(kFinalizeThreads) k_finalize(BatchArgs a, int use_global) {
    const int s = (int)blockIdx.x;
    if (s >= a.n_streams) return;
    const Coop co{(int)threadIdx.x, (int)blockDim.x};
    float* ws = use_global ? a.fin_scratch + (size_t)s * finalize_workspace_floats(a.n_rows, a.n_pad) : fin_smem;
    body_finalize(a, s, co, ws);
}

// EQ magnitude response (dsp/eq.rs:514-527 over dsp/biquad.rs:184-205): one thread per (set, frequency).
__global__ void k_eq_response(const double* __restrict__ coeffs /* [set][40][5] */,
                              const int* __restrict__ n_sections /* [set][10] sections per band */,
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
        for (int k = 0; k < ns; ++k) {
            const double* c = coeffs + ((size_t)set * kMaxSections + band * 4 + k) * 5;
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
__global__ void k_synth(float* out, size_t n_per, int n_passages, int kind, double sample_rate) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = n_per * (size_t)n_passages;
    if (i >= total) return;
    const int p = (int)(i / n_per);
    const size_t n = i % n_per;
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
        const float h = 0.30f * sinf(6.2831853f * f0 * t) + 0.14f * sinf(6.2831853f * 2.0f * f0 * t) +
                        0.10f * sinf(6.2831853f * 3.0f * f0 * t) + 0.08f * sinf(6.2831853f * 2700.0f * t);
        const int gate = ((int)(t / 0.12f) % 5) == 2;
        v = 0.70f * (env * h + (gate ? 0.30f * sinf(6.2831853f * 7200.0f * t) : 0.0f) + u * 0.0126f);
    }
    out[i] = v;
}

// Issue-rate microbenchmarks (bench.py's roofline denominators for the recurrence / FIR stages):
// 8 independent dependent chains per thread of DMUL+DADD (kind 0; -fmad=false keeps them apart, the
// same instruction mix the biquads issue) or FFMA (kind 1).
__global__ void __launch_bounds__(256) k_issue_peak(int kind, int iters, double* sink) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (kind == 0) {
        double a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 1.0 + 1e-9 * (tid + j);
        const double m = 1.0 - 1e-12 * (tid & 7), c = 1e-13;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = a[j] * m + c;
        }
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) r += a[j];
        if (r == 12345.678) sink[0] = r;
    } else {
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 1.0f + 1e-6f * (tid + j);
        const float m = 1.0f - 1e-7f * (tid & 7), c = 1e-8f;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __fmaf_rn(a[j], m, c);
        }
        float r = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) r += a[j];
        if (r == 12345.678f) sink[0] = r;
    }
}

// Self-test of afsim_math.h: the custom routines against the CUDA math library / division on hashed arguments;
// counts[k] = results that differ in any bit (k: 0 log10, 1 exp10, 2 x/20, 3 x/40, 4 x/3.75, 5 a/b through AfDivisor).
__global__ void __launch_bounds__(256) k_selftest_math(unsigned long long n, unsigned long long* counts) {
    unsigned long long bad[6] = {0, 0, 0, 0, 0, 0};  // [4] also covers the de-esser's norm_range divisors
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        uint64_t z = 0x9e3779b97f4a7c15ull * (i + 1);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        z = z ^ (z >> 31);
        const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
        const double v = (double)((z * 0xd1342543de82ef95ull) >> 11) * (1.0 / 9007199254740992.0);
        // log10: magnitudes the chain sees (1e-10 floors .. a few), plus the whole normal range now and then
        const int wide = (int)(z & 15) == 0;
        const double lx = wide ? __longlong_as_double((long long)((z >> 1) & 0x7fefffffffffffffull) | 0x0010000000000000ll)
                               : exp2(-40.0 + 44.0 * u) * (1.0 + v);
        // exp10 / divisions: dB-sized arguments, both signs, plus tiny and huge ones now and then
        const double ex = wide ? (u - 0.5) * 700.0 : (u - 0.5) * 24.0;
        const double dx = wide ? __longlong_as_double((long long)(z & 0xffefffffffffffffull)) : (u - 0.5) * 240.0 * v;
        bad[0] += __double_as_longlong(af_log10(lx)) != __double_as_longlong(log10(lx));
        bad[1] += __double_as_longlong(af_exp10_fast(ex)) != __double_as_longlong(exp10(ex));
        bad[2] += __double_as_longlong(af_div_const(dx, 20.0, 0.05)) != __double_as_longlong(dx / 20.0);
        bad[3] += __double_as_longlong(af_div_const(dx, 40.0, 0.025)) != __double_as_longlong(dx / 40.0);
        bad[4] += __double_as_longlong(af_div_const(dx, 3.75, 1.0 / 3.75)) != __double_as_longlong(dx / 3.75);
        bad[4] += __double_as_longlong(af_div_const(dx, 10.0 - 1.5, 1.0 / (10.0 - 1.5))) != __double_as_longlong(dx / (10.0 - 1.5));
        bad[4] += __double_as_longlong(af_div_const(dx, -24.0 - -62.0, 1.0 / (-24.0 - -62.0))) != __double_as_longlong(dx / (-24.0 - -62.0));
        bad[4] += __double_as_longlong(af_div_const(dx, -34.0 - -58.0, 1.0 / (-34.0 - -58.0))) != __double_as_longlong(dx / (-34.0 - -58.0));
        bad[4] += __double_as_longlong(af_div_const(dx, 0.68 - 0.34, 1.0 / (0.68 - 0.34))) != __double_as_longlong(dx / (0.68 - 0.34));
        // prepared divisor: biquad-sized operands, and arbitrary finite ones now and then
        const double num = wide ? dx : (v - 0.5) * 4.0;
        const double den = wide ? __longlong_as_double((long long)((z * 0x2545f4914f6cdd1dull) & 0xffefffffffffffffull)) : 0.25 + 3.0 * u;
        bad[5] += __double_as_longlong(af_div(num, af_divisor(den))) != __double_as_longlong(num / den);
    }
    for (int k = 0; k < 6; ++k)
        if (bad[k]) atomicAdd(counts + k, bad[k]);
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
cudaError_t launch_selftest_math(unsigned long long n, unsigned long long* counts, cudaStream_t st) {
    k_selftest_math<<<148 * 4, 256, 0, st>>>(n, counts);
    return cudaGetLastError();
}
cudaError_t launch_issue_peak(int kind, int iters, int blocks, double* sink, cudaStream_t st) {
    k_issue_peak<<<blocks, 256, 0, st>>>(kind, iters, sink);
    return cudaGetLastError();
}

static inline dim3 stream_grid(const BatchArgs& a, int block) { return dim3((unsigned)((a.n_streams + block - 1) / block)); }

// Few-stream batches use 32-thread blocks so that the warps spread over all SMs.
static inline int pick_block(const BatchArgs& a) { return a.n_streams <= 148 * 4 * 32 ? 32 : 128; }

cudaError_t launch_expand_deesser(const BatchArgs& a, cudaStream_t st) {
    const int b = pick_block(a);
    k_expand_deesser<<<stream_grid(a, b), b, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_input(const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st) {
    const int b = pick_block(a);
    if (input_uses_cleanup(a) && a.n_streams <= kWarpPerStreamMax && a.n_streams > 0)
        k_input_cleanup_warp<<<(unsigned)a.n_streams, 32, 0, st>>>(a, ck);  // few streams: one warp each, a lane per hum bin
    else if (input_uses_cleanup(a))
        k_input_cleanup<<<stream_grid(a, b), b, 0, st>>>(a, ck);
    else
        k_input<<<stream_grid(a, b), b, 0, st>>>(a, ck);
    return cudaGetLastError();
}
cudaError_t launch_input_fanout(const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st) {
    const int n_groups = (a.in_det || a.in_de[0]) ? 1 : (ck.len + kFanoutGroup - 1) / kFanoutGroup;  // rows and statistics only
    const dim3 grid((unsigned)((a.n_streams + 31) / 32), (unsigned)((n_groups + kMapWarps - 1) / kMapWarps));
    k_input_fanout<<<grid, 32 * kMapWarps, 0, st>>>(a, ck);
    return cudaGetLastError();
}
cudaError_t launch_eq(const BatchArgs& a, const ChunkArgs& ck, int first, int k, cudaStream_t st) {
    const int b = pick_block(a);
    if (k == 10)
        k_eq<10><<<stream_grid(a, b), b, 0, st>>>(a, ck, first);
    else if (k == 5)
        k_eq<5><<<stream_grid(a, b), b, 0, st>>>(a, ck, first);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}
cudaError_t launch_compressor(const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st) {
    const int b = pick_block(a);
    if (a.in_det)
        k_compressor_shared<<<stream_grid(a, b), b, 0, st>>>(a, ck);
    else
        k_compressor<<<stream_grid(a, b), b, 0, st>>>(a, ck);
    return cudaGetLastError();
}
cudaError_t launch_limiter(const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st) {
    const int b = pick_block(a);
    k_limiter<<<stream_grid(a, b), b, 0, st>>>(a, ck);
    return cudaGetLastError();
}
cudaError_t launch_output(const BatchArgs& a, const ChunkArgs& ck, bool limiter, cudaStream_t st) {
    const int b = pick_block(a);
    if (limiter)
        k_output<true><<<stream_grid(a, b), b, 0, st>>>(a, ck);
    else
        k_output<false><<<stream_grid(a, b), b, 0, st>>>(a, ck);
    return cudaGetLastError();
}
cudaError_t launch_input_true_peak(const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st) {
    const int b = pick_block(a);
    k_input_true_peak<<<stream_grid(a, b), b, 0, st>>>(a, ck);
    return cudaGetLastError();
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

cudaError_t launch_split(SplitOp op, const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st) {
    // serial kernels: one warp per block with a shared-memory staging area for few-stream batches; plain
    // 128-thread blocks reading global memory directly when the batch itself fills the GPU
    const int rb = a.stage_inputs ? kRBlock : 128;
    const dim3 rgrid = stream_grid(a, rb);
    const size_t lane_bytes = (op == SP_DE_RC || op == SP_DE_RC1A) ? kDeRc1StagingBytesPerLane
                              : (op == SP_DE_RC1C ? kDeRc1cStagingBytesPerLane
                                                  : (op == SP_DE_RC3 ? kDeRc3StagingBytesPerLane : kStagingBytesPerLane));
    const size_t rsm = a.stage_inputs ? lane_bytes * kRBlock : 0;
    const int mb = 32 * kMapWarps;
    const int group = (op == SP_COMP_M2 || op == SP_COMP_M4 || op == SP_COMP_M6) ? kCompMapGroup
                      : (op == SP_LIM_M ? kLimGroup : (op == SP_DE_MB ? kDeMapGroup : (op == SP_DE_MC2 ? kDeRebuildGroup : (op == SP_DE_MC1B ? kDeTargetGroup : kGroup))));
    const int n_groups = (ck.len + group - 1) / group;
    // Map kernels are FP64 / FP32 issue bound and need only a few warps per SM to saturate the pipe; capping
    // the grid (blocks loop over sample groups) leaves issue slots for the co-resident serial kernels of the
    // other wavefront stages.  AFSIM_MAP_BLOCKS_PER_SM = 0 removes the cap.
    const int blocks_per_sm = a.map_blocks_per_sm;  // read with the other knobs when the sweep is built (afsim_api.cu)
    const unsigned gx = (unsigned)((a.n_streams + 31) / 32);
    unsigned gy = (unsigned)((n_groups + kMapWarps - 1) / kMapWarps);
    // The FP32 FIR maps and the limiter window map have so much instruction-level parallelism per warp that a few
    // resident warps saturate the FMA issue; every further resident warp only takes issue slots from the serial
    // kernels' warps on the same scheduler (round-robin among ready warps).  Their grids are capped (blocks loop
    // over sample groups).  AFSIM_FIR_BLOCKS_PER_SM overrides the cap (0 = none).
    const int fir_blocks_env = a.fir_blocks_per_sm;
    const bool fir_like = op == SP_TP_FIR_IN || op == SP_TP_FIR_OUT || op == SP_LIM_M;
    int cap_per_sm = blocks_per_sm;
    if (fir_like) cap_per_sm = fir_blocks_env >= 0 ? fir_blocks_env : kFirBlocksPerSm;
    if (cap_per_sm > 0) {
        const unsigned cap = (unsigned)((cap_per_sm * sm_count() + (int)gx - 1) / (int)gx);
        if (gy > cap) gy = cap < 1 ? 1 : cap;
    }
    const dim3 mgrid(gx, gy);
    switch (op) {
        case SP_COMP_R1: k_comp_r1<<<rgrid, rb, rsm, st>>>(a, ck); break;
        case SP_COMP_M2: k_comp_m2<<<mgrid, mb, 0, st>>>(a, ck); break;
        case SP_COMP_R3: k_comp_r3<<<rgrid, rb, rsm, st>>>(a, ck); break;
        case SP_COMP_M4: k_comp_m4<<<mgrid, mb, 0, st>>>(a, ck); break;
        case SP_COMP_R5: k_comp_r5<<<rgrid, rb, rsm, st>>>(a, ck); break;
        case SP_COMP_M6: k_comp_m6<<<mgrid, mb, 0, st>>>(a, ck); break;
        case SP_LIM_M: k_lim_m<<<mgrid, mb, 0, st>>>(a, ck); break;
        case SP_LIM_R: k_lim_r<<<rgrid, rb, rsm, st>>>(a, ck); break;
        case SP_TP_FIR_IN: k_tp_fir_in<<<mgrid, mb, 0, st>>>(a, ck); break;
        case SP_TP_R: k_tp_r<<<rgrid, rb, rsm, st>>>(a, ck); break;
        case SP_TP_FIR_OUT: k_tp_fir_out<<<mgrid, mb, 0, st>>>(a, ck); break;
        case SP_DE_RA:
            if (a.stage_inputs)
                k_de_ra<<<rgrid, rb, rsm, st>>>(a, ck);
            else
                k_de_ra_direct<<<rgrid, rb, 0, st>>>(a, ck);
            break;
        case SP_DE_MB: k_de_mb<<<mgrid, mb, 0, st>>>(a, ck); break;
        case SP_DE_RC:
            if (a.stage_inputs)
                k_de_rc<<<rgrid, rb, rsm, st>>>(a, ck);
            else
                k_de_rc_direct<<<rgrid, rb, 0, st>>>(a, ck);
            break;
        case SP_DE_RC1A:
            if (a.stage_inputs)
                k_de_rc1a<<<rgrid, rb, rsm, st>>>(a, ck);
            else
                k_de_rc1a_direct<<<rgrid, rb, 0, st>>>(a, ck);
            break;
        case SP_DE_MC1B: k_de_mc1b<<<mgrid, mb, 0, st>>>(a, ck); break;
        case SP_DE_RC1C:
            if (a.stage_inputs)
                k_de_rc1c<<<rgrid, rb, rsm, st>>>(a, ck);
            else
                k_de_rc1c_direct<<<rgrid, rb, 0, st>>>(a, ck);
            break;
        case SP_COMP_R7: k_comp_r7<<<rgrid, rb, rsm, st>>>(a, ck); break;
        case SP_DE_MC2: k_de_mc2<<<mgrid, mb, 0, st>>>(a, ck); break;
        case SP_DE_RC3:
            if (a.stage_inputs) {
                k_de_rc3<<<rgrid, rb, rsm, st>>>(a, ck);  // > 48 KB of staging: opted in by configure_kernels()
            } else {
                k_de_rc3_direct<<<rgrid, rb, 0, st>>>(a, ck);
            }
            break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// Per-device kernel attributes (called by afsim_create with the device current).
cudaError_t configure_kernels() {
    return cudaFuncSetAttribute(k_de_rc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kDeRc3StagingBytesPerLane * kRBlock));
}

size_t finalize_workspace_bytes(int n_rows, int n_pad) { return finalize_workspace_floats(n_rows, n_pad) * sizeof(float); }

cudaError_t launch_finalize(const BatchArgs& a, cudaStream_t st) {
    if (a.n_streams <= 0) return cudaSuccess;
    const size_t bytes = finalize_workspace_bytes(a.n_rows, a.n_pad);
    const int use_global = bytes > kFinalizeSmemLimit ? 1 : 0;
    if (use_global && !a.fin_scratch) return cudaErrorInvalidValue;
    size_t smem = use_global ? 0 : bytes;
    if (smem > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
    }
    k_finalize<<<(unsigned)a.n_streams, kFinalizeThreads, smem, st>>>(a, use_global);
    return cudaGetLastError();
}

cudaError_t launch_eq_response(const double* coeffs, const int* n_sections, const double* freqs, int n_freqs,
                               int n_sets, double fs, double* out, cudaStream_t st) {
    const int total = n_freqs * n_sets;
    if (total <= 0) return cudaSuccess;
    k_eq_response<<<(total + 127) / 128, 128, 0, st>>>(coeffs, n_sections, freqs, n_freqs, n_sets, fs, out);
    return cudaGetLastError();
}

cudaError_t launch_synth(float* out, size_t n_per, int n_passages, int kind, double fs, cudaStream_t st) {
    const size_t total = n_per * (size_t)n_passages;
    if (total == 0) return cudaSuccess;
    k_synth<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, n_per, n_passages, kind, fs);
    return cudaGetLastError();
}

}  // namespace afsim
