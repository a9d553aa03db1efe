// afsim_eqscan.cu -- sm_100a kernels of the time-parallel EQ render of one long passage (afsim_eqscan.h) and the
// parallel input / output statistics of simulate_eq_v2 (lib.rs:231-262: sample peak, RMS, 4x true peak).
#include <cuda_runtime.h>

#include "afsim_eqscan.h"
#include "afsim_kernels.h"

namespace afsim {

__constant__ float c_fir_scan[4][32] = {
#include "true_peak_fir.inc"
};

// in[seg * L + i] -> xt[i * P + seg] (zero padded): neighbouring threads then walk neighbouring segments coalesced
__global__ void __launch_bounds__(256) k_eqscan_transpose_in(const float* __restrict__ in, size_t n, float* __restrict__ xt,
                                                             size_t n_seg, int len) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_seg * (size_t)len) return;
    const size_t seg = idx % n_seg, i = idx / n_seg, src = seg * (size_t)len + i;
    xt[idx] = src < n ? in[src] : 0.0f;
}
__global__ void __launch_bounds__(256) k_eqscan_transpose_out(const float* __restrict__ xt, size_t n, size_t n_seg, int len,
                                                              float* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    out[idx] = xt[(idx % (size_t)len) * n_seg + idx / (size_t)len];
}

template <bool APPLY, bool LOCAL>
__global__ void __launch_bounds__(128) k_eqscan_pass(float* xt, size_t n_seg, int len, Bq cur, const double* __restrict__ init,
                                                     Bq nxt, double* __restrict__ end) {
    const size_t seg = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_seg) return;
    double e1 = 0.0, e2 = 0.0;
    eqscan_segment<APPLY, LOCAL>(xt, n_seg, seg, len, cur, APPLY ? init[seg] : 0.0, APPLY ? init[n_seg + seg] : 0.0, nxt, &e1, &e2);
    if (LOCAL) {
        end[seg] = e1;
        end[n_seg + seg] = e2;
    }
}

__device__ __forceinline__ Affine2 shfl_up_affine(const Affine2& a, int delta) {
    Affine2 r;
    r.m00 = __shfl_up_sync(0xffffffffu, a.m00, delta);
    r.m01 = __shfl_up_sync(0xffffffffu, a.m01, delta);
    r.m10 = __shfl_up_sync(0xffffffffu, a.m10, delta);
    r.m11 = __shfl_up_sync(0xffffffffu, a.m11, delta);
    r.v0 = __shfl_up_sync(0xffffffffu, a.v0, delta);
    r.v1 = __shfl_up_sync(0xffffffffu, a.v1, delta);
    return r;
}

// Start state of every segment from the zero-state end states: one block; each thread owns `per` consecutive
// segments, the per-thread composites are scanned with warp shuffles and combined across warps in shared memory.
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) k_eqscan_scan(Bq c, int log2_len, size_t n_seg, const double* __restrict__ end,
                                                              double* __restrict__ init) {
    __shared__ Affine2 warp_total[kScanThreads / 32];
    const int t = (int)threadIdx.x, lane = t & 31, warp = t >> 5;
    double m[4];
    biquad_transition_power(c, log2_len, m);
    const size_t per = (n_seg + kScanThreads - 1) / kScanThreads;
    const size_t k0 = (size_t)t * per, k1 = k0 + per < n_seg ? k0 + per : n_seg;
    Affine2 comp = affine_identity();
    constexpr int kPre = 8;  // this thread's end states, loaded together (per <= 8 up to 8192 segments)
    double pe1[kPre], pe2[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
        const size_t k = k0 + j;
        pe1[j] = k < k1 ? end[k] : 0.0;
        pe2[j] = k < k1 ? end[n_seg + k] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < kPre; ++j)
        if (k0 + j < k1) comp = affine_then(comp, Affine2{m[0], m[1], m[2], m[3], pe1[j], pe2[j]});
    for (size_t k = k0 + kPre; k < k1; ++k) comp = affine_then(comp, Affine2{m[0], m[1], m[2], m[3], end[k], end[n_seg + k]});
    // inclusive Kogge-Stone scan inside the warp
    Affine2 incl = comp;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Affine2 prev = shfl_up_affine(incl, d);
        if (lane >= d) incl = affine_then(prev, incl);
    }
    if (lane == 31) warp_total[warp] = incl;
    __syncthreads();
    if (warp == 0) {  // scan of the 32 warp totals
        Affine2 w = warp_total[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Affine2 prev = shfl_up_affine(w, d);
            if (lane >= d) w = affine_then(prev, w);
        }
        warp_total[lane] = w;
    }
    __syncthreads();
    // exclusive prefix of this thread = (all earlier warps) then (earlier lanes of this warp)
    Affine2 excl = shfl_up_affine(incl, 1);
    if (lane == 0) excl = affine_identity();
    if (warp > 0) excl = affine_then(warp_total[warp - 1], excl);
    double s1 = excl.v0, s2 = excl.v1;  // the passage starts from a zero state
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
        const size_t k = k0 + j;
        if (k < k1) {
            init[k] = s1;
            init[n_seg + k] = s2;
            const double t1 = m[0] * s1 + m[1] * s2 + pe1[j];
            const double t2 = m[2] * s1 + m[3] * s2 + pe2[j];
            s1 = t1;
            s2 = t2;
        }
    }
    for (size_t k = k0 + kPre; k < k1; ++k) {
        init[k] = s1;
        init[n_seg + k] = s2;
        const double t1 = m[0] * s1 + m[1] * s2 + end[k];
        const double t2 = m[2] * s1 + m[3] * s2 + end[n_seg + k];
        s1 = t1;
        s2 = t2;
    }
}

// ---- statistics of a plain array: |x| peak, 4x true peak, sum of squares (deterministic tree) ------------------
struct PlainStats {
    double sum;
    float peak, tp_peak;
    unsigned non_finite, pad;
};
constexpr int kStatThreads = 256;
__global__ void __launch_bounds__(kStatThreads) k_plain_stats(const float* __restrict__ x, size_t n, PlainStats* __restrict__ partial) {
    __shared__ double s_sum[kStatThreads];
    __shared__ float s_peak[kStatThreads], s_tp[kStatThreads];
    __shared__ unsigned s_bad[kStatThreads];
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t t0 = g * kFirChunk;
    double sum = 0.0;
    float peak = 0.0f, tp = 0.0f;
    unsigned bad = 0;
    if (t0 < n) {
        float win[kFirWin];
#pragma unroll
        for (int i = 0; i < kFirWin; ++i) {
            const long long idx = (long long)t0 + i - 31;
            const float v = (idx >= 0 && (size_t)idx < n) ? x[idx] : 0.0f;
            win[i] = af_finite(v) ? v : 0.0f;  // dsp/true_peak.rs:211 sanitise
        }
        float pk[kFirChunk];
        fir8_peaks(win, c_fir_scan, pk);
#pragma unroll
        for (int j = 0; j < kFirChunk; ++j) {
            if (t0 + j < n) {
                const float s = x[t0 + j];
                peak = fmaxf(peak, fabsf(s));
                tp = fmaxf(tp, pk[j]);
                sum += (double)s * (double)s;
                bad |= af_finite(s) ? 0u : 1u;
            }
        }
    }
    const int t = (int)threadIdx.x;
    s_sum[t] = sum;
    s_peak[t] = peak;
    s_tp[t] = tp;
    s_bad[t] = bad;
    __syncthreads();
    for (int w = kStatThreads / 2; w > 0; w >>= 1) {
        if (t < w) {
            s_sum[t] += s_sum[t + w];
            s_peak[t] = fmaxf(s_peak[t], s_peak[t + w]);
            s_tp[t] = fmaxf(s_tp[t], s_tp[t + w]);
            s_bad[t] |= s_bad[t + w];
        }
        __syncthreads();
    }
    if (t == 0) partial[blockIdx.x] = PlainStats{s_sum[0], s_peak[0], s_tp[0], s_bad[0], 0};
}
__global__ void __launch_bounds__(kStatThreads) k_plain_stats_final(const PlainStats* __restrict__ partial, int n_partial,
                                                                    PlainStats* __restrict__ out) {
    __shared__ double s_sum[kStatThreads];
    __shared__ float s_peak[kStatThreads], s_tp[kStatThreads];
    __shared__ unsigned s_bad[kStatThreads];
    const int t = (int)threadIdx.x;
    double sum = 0.0;
    float peak = 0.0f, tp = 0.0f;
    unsigned bad = 0;
    for (int i = t; i < n_partial; i += kStatThreads) {
        sum += partial[i].sum;
        peak = fmaxf(peak, partial[i].peak);
        tp = fmaxf(tp, partial[i].tp_peak);
        bad |= partial[i].non_finite;
    }
    s_sum[t] = sum;
    s_peak[t] = peak;
    s_tp[t] = tp;
    s_bad[t] = bad;
    __syncthreads();
    for (int w = kStatThreads / 2; w > 0; w >>= 1) {
        if (t < w) {
            s_sum[t] += s_sum[t + w];
            s_peak[t] = fmaxf(s_peak[t], s_peak[t + w]);
            s_tp[t] = fmaxf(s_tp[t], s_tp[t + w]);
            s_bad[t] |= s_bad[t + w];
        }
        __syncthreads();
    }
    if (t == 0) *out = PlainStats{s_sum[0], s_peak[0], s_tp[0], s_bad[0], 0};
}

// ---- launcher -------------------------------------------------------------------------------------------------------
size_t eqscan_stats_partials(size_t n) {
    const size_t groups = (n + kFirChunk - 1) / kFirChunk;
    return (groups + kStatThreads - 1) / kStatThreads;
}

cudaError_t launch_plain_stats(const float* x, size_t n, void* partial, void* out, cudaStream_t st) {
    const size_t blocks = eqscan_stats_partials(n);
    if (blocks == 0) return cudaMemsetAsync(out, 0, sizeof(PlainStats), st);
    k_plain_stats<<<(unsigned)blocks, kStatThreads, 0, st>>>(x, n, static_cast<PlainStats*>(partial));
    k_plain_stats_final<<<1, kStatThreads, 0, st>>>(static_cast<const PlainStats*>(partial), (int)blocks, static_cast<PlainStats*>(out));
    return cudaGetLastError();
}

// d_in -> d_out through `n_sections` DF2T sections (coeffs[j] = b0 b1 b2 a1 a2); xt: n_seg * len floats;
// seg_state: 4 * n_seg doubles (end states, start states).  Returns the number of kernels launched in *launches.
cudaError_t launch_eqscan(const float* d_in, float* d_out, size_t n, const double (*coeffs)[5], int n_sections, int log2_len,
                          float* xt, double* seg_state, int* launches, cudaStream_t st) {
    const int len = 1 << log2_len;
    const size_t n_seg = (n + (size_t)len - 1) / (size_t)len;
    int count = 0;
    if (n == 0) {
        *launches = 0;
        return cudaSuccess;
    }
    if (n_sections == 0) {
        *launches = 0;
        return cudaMemcpyAsync(d_out, d_in, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
    }
    double* end = seg_state;
    double* init = seg_state + 2 * n_seg;
    const unsigned tb = (unsigned)((n_seg * (size_t)len + 255) / 256);
    k_eqscan_transpose_in<<<tb, 256, 0, st>>>(d_in, n, xt, n_seg, len);
    ++count;
    const unsigned pb = (unsigned)((n_seg + 127) / 128);
    Bq first = bq_from(coeffs[0]);
    k_eqscan_pass<false, true><<<pb, 128, 0, st>>>(xt, n_seg, len, first, nullptr, first, end);
    ++count;
    for (int j = 0; j < n_sections; ++j) {
        const Bq cur = bq_from(coeffs[j]);
        k_eqscan_scan<<<1, kScanThreads, 0, st>>>(cur, log2_len, n_seg, end, init);
        ++count;
        if (j + 1 < n_sections) {
            const Bq nxt = bq_from(coeffs[j + 1]);
            k_eqscan_pass<true, true><<<pb, 128, 0, st>>>(xt, n_seg, len, cur, init, nxt, end);
        } else {
            k_eqscan_pass<true, false><<<pb, 128, 0, st>>>(xt, n_seg, len, cur, init, cur, end);
        }
        ++count;
    }
    k_eqscan_transpose_out<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(xt, n, n_seg, len, d_out);
    ++count;
    *launches = count;
    return cudaGetLastError();
}

}  // namespace afsim
