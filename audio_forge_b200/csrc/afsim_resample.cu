// Product resampler simulator: host planner + device kernel (see afsim_resample.h).
//
// The reference streams the signal through rubato 0.14's SincFixedIn<f64> in blocks (resampling.rs:228-259).  What a
// frame reads does not depend on the blocks -- only WHERE the frames sit does (the block loop accumulates the position
// with one f64 addition per frame and re-bases it by -chunk per block) -- so the planner walks the block loop once on the
// host, addition for addition, and records per frame the input window, the phase and the cubic's abscissa; every frame of
// every stream is then an independent 4 x sinc_len dot product, rendered by one warp with coalesced phase-table loads.
#include "afsim_resample.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace afsim {

namespace {

const char* window_name(int w) {
    switch (w) {
        case AF_WINDOW_BLACKMAN_HARRIS: return "blackman_harris";
        case AF_WINDOW_BLACKMAN_HARRIS2: return "blackman_harris_squared";
        case AF_WINDOW_BLACKMAN: return "blackman";
        case AF_WINDOW_BLACKMAN2: return "blackman_squared";
        case AF_WINDOW_HANN: return "hann";
        case AF_WINDOW_HANN2: return "hann_squared";
        default: return nullptr;
    }
}

// rubato 0.14 `calculate_cutoff(sinc_len, window)` as f32, for the configurations the reference ships / evaluates
// (resampling.rs:131-138, python/tools/evaluate_resampler_quality.py).  The crate is not vendored in the reference tree;
// each value is the one f32 that reproduces the reference's published measurements of the real crate
// (evaluation/resampler-quality-report.json; tools/fit_resampler_cutoff.py: the adjacent f32 values miss them by ~1e-5 dB).
bool pinned_cutoff(int sinc_len, int window, float* out) {
    if (sinc_len == 128 && window == AF_WINDOW_BLACKMAN) return *out = 0.9527542591094971f, true;
    if (sinc_len == 128 && window == AF_WINDOW_BLACKMAN_HARRIS2) return *out = 0.8947277069091797f, true;
    if (sinc_len == 256 && window == AF_WINDOW_BLACKMAN_HARRIS2) return *out = 0.9470546841621399f, true;
    return false;
}

double window_value(int window, double x, double n) {  // periodic cosine-sum windows (rubato windows.rs)
    const double pi = 3.14159265358979323846;
    double w;
    switch (window) {
        case AF_WINDOW_BLACKMAN_HARRIS:
        case AF_WINDOW_BLACKMAN_HARRIS2:
            w = 0.35875 - 0.48829 * std::cos(2.0 * pi * x / n) + 0.14128 * std::cos(4.0 * pi * x / n) - 0.01168 * std::cos(6.0 * pi * x / n);
            break;
        case AF_WINDOW_BLACKMAN:
        case AF_WINDOW_BLACKMAN2:
            w = 0.42 - 0.5 * std::cos(2.0 * pi * x / n) + 0.08 * std::cos(4.0 * pi * x / n);
            break;
        default:
            w = 0.5 - 0.5 * std::cos(2.0 * pi * x / n);
            break;
    }
    const bool squared = window == AF_WINDOW_BLACKMAN_HARRIS2 || window == AF_WINDOW_BLACKMAN2 || window == AF_WINDOW_HANN2;
    return squared ? w * w : w;
}

void build_table(int sinc_len, double cutoff, int window, std::vector<double>* table) {  // rubato sinc.rs make_sincs
    const double pi = 3.14159265358979323846;
    const int factor = kResamplePhases;
    const size_t tot = static_cast<size_t>(sinc_len) * factor;
    std::vector<double> y(tot);
    double sum = 0.0;
    for (size_t x = 0; x < tot; ++x) {
        const double arg = (static_cast<double>(x) - static_cast<double>(tot / 2)) * cutoff / static_cast<double>(factor);
        const double s = arg == 0.0 ? 1.0 : std::sin(arg * pi) / (arg * pi);
        const double val = window_value(window, static_cast<double>(x), static_cast<double>(tot)) * s;
        sum += val;
        y[x] = val;
    }
    sum /= static_cast<double>(factor);
    table->resize(tot);
    for (int p = 0; p < sinc_len; ++p)
        for (int n = 0; n < factor; ++n) (*table)[static_cast<size_t>(factor - n - 1) * sinc_len + p] = y[static_cast<size_t>(factor) * p + n] / sum;
}

}  // namespace

int plan_resampler(const AfResamplerSpec& spec, size_t n_in, bool with_frames, bool with_table, ResamplePlan* plan, std::string* msg) {
    if (spec.input_rate == 0 || spec.output_rate == 0) return *msg = "sample rates must be positive", AFSIM_INVALID_ARGUMENT;
    if (spec.chunk_size < 1 || spec.chunk_size > 1024) return *msg = "chunk_size must be between 1 and 1024", AFSIM_INVALID_ARGUMENT;
    const uint32_t sl = spec.sinc_len;
    if (sl < 32 || sl > 2048 || (sl & (sl - 1)) != 0)
        return *msg = "sinc_len must be a power of two between 32 and 2048", AFSIM_INVALID_ARGUMENT;
    if (!window_name(spec.window)) return *msg = "unsupported resampler window", AFSIM_INVALID_ARGUMENT;
    float cutoff32 = 0.0f;
    if (!pinned_cutoff(static_cast<int>(sl), spec.window, &cutoff32)) {
        char buf[256];
        std::snprintf(buf, sizeof buf,
                      "no pinned rubato calculate_cutoff value for sinc_len %u with window %s (pinned: 128 blackman, 128 / 256 blackman_harris_squared)",
                      sl, window_name(spec.window));
        return *msg = buf, AFSIM_UNSUPPORTED;
    }
    const double ratio = static_cast<double>(spec.output_rate) / static_cast<double>(spec.input_rate);
    if (ratio < 1.0) cutoff32 = cutoff32 * static_cast<float>(ratio);  // make_interpolator, f32 product
    plan->spec = spec;
    plan->n_in = n_in;
    plan->cutoff = static_cast<double>(cutoff32);
    const double expected_f = std::round(static_cast<double>(n_in) * static_cast<double>(spec.output_rate) / static_cast<double>(spec.input_rate));
    if (expected_f > 2.0e9 || n_in > 0x7ff00000ull) return *msg = "signal too long", AFSIM_INVALID_ARGUMENT;  // frame windows are int32
    plan->shape.expected_frames = static_cast<uint64_t>(expected_f);
    plan->shape.delay = static_cast<uint32_t>(static_cast<double>(sl / 2) * ratio);
    // the block loop of SincFixedIn::process_into_buffer, frames only
    const double t_ratio = 1.0 / ratio;
    const double chunk = static_cast<double>(spec.chunk_size);
    const double end_idx = static_cast<double>(static_cast<long long>(spec.chunk_size) - (static_cast<long long>(sl) + 1) -
                                               static_cast<long long>(std::ceil(t_ratio)));
    const uint64_t blocks_in = (n_in + spec.chunk_size - 1) / spec.chunk_size;
    const uint64_t need = plan->shape.expected_frames + plan->shape.delay;
    double last = -static_cast<double>(sl / 2);
    uint64_t produced = 0, b = 0;
    plan->frames.clear();
    if (with_frames) plan->frames.reserve(static_cast<size_t>(need + 2 * spec.chunk_size * ratio + 16));
    while (b < blocks_in || produced < need) {
        double idx = last;
        uint64_t k = 0;
        while (idx < end_idx) {
            idx += t_ratio;
            if (with_frames) {
                const double fl = std::floor(idx);
                const double scaled = idx * static_cast<double>(kResamplePhases);
                ResampleFrame f;
                f.base = static_cast<int32_t>(static_cast<long long>(b) * spec.chunk_size + static_cast<long long>(fl));
                f.sub = static_cast<int32_t>(std::floor((idx - fl) * static_cast<double>(kResamplePhases)));
                f.frac = scaled - std::floor(scaled);
                plan->frames.push_back(f);
            }
            ++k;
        }
        if (k == 0 && b >= blocks_in)
            return *msg = "resampler flush produced no frames before reaching the expected length", AFSIM_INVALID_ARGUMENT;
        produced += k;
        if (produced > 0x7ff00000ull) return *msg = "signal too long", AFSIM_INVALID_ARGUMENT;
        last = idx - chunk;
        ++b;
    }
    plan->shape.frames = produced;
    plan->shape.blocks = static_cast<uint32_t>(b);
    plan->max_span = 0;
    if (with_frames) {
        for (size_t f0 = 0; f0 < plan->frames.size(); f0 += kResampleFramesPerBlock) {
            const size_t f1 = std::min(plan->frames.size(), f0 + kResampleFramesPerBlock) - 1;
            const int span = plan->frames[f1].base - plan->frames[f0].base + static_cast<int>(sl) + 3;
            if (span > plan->max_span) plan->max_span = span;
        }
        if (plan->max_span > 6000) return *msg = "rate ratio too extreme for this build (input span of one frame group exceeds shared memory)", AFSIM_UNSUPPORTED;
    }
    if (with_table) build_table(static_cast<int>(sl), plan->cutoff, spec.window, &plan->table);
    return AFSIM_OK;
}

// One CTA = 128 consecutive frames of S streams: their input spans go through shared memory once; one warp per frame, lane l
// owns taps l, l + 32, ... of the four phases (phase-table rows read as whole 256-byte lines, each value used for all S
// streams), the cubic is applied to the lane's partial sums (it is linear in the four points) and one reduction over the
// lanes (pairs at distance 1, 2, 4, 8, 16 -- the same tree for every S, so a batched render equals single renders bit for
// bit) finishes the frame.  The kernel is bound by L1 bandwidth on the phase table (4 x sinc_len x 8 bytes per frame, every
// frame another set of rows), which is why S streams share each table load.
template <int S>
__device__ __forceinline__ void reduce_and_keep(double (&v)[S], int lane, int k, double* mine);

template <>
__device__ __forceinline__ void reduce_and_keep<1>(double (&v)[1], int lane, int k, double* mine) {
    double t = v[0];
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
    if (lane == k) *mine = t;  // frame k of the warp's 32 lands in lane k: one coalesced store per warp
}

template <>
__device__ __forceinline__ void reduce_and_keep<4>(double (&v)[4], int lane, int k, double* mine) {
    // distance 1: even lanes go on with streams 0 / 1, odd lanes with 2 / 3; distance 2: bit 1 picks between the two
    const bool b0 = lane & 1, b1 = lane & 2;
    const double r0 = __shfl_xor_sync(0xffffffffu, b0 ? v[0] : v[2], 1);
    const double r1 = __shfl_xor_sync(0xffffffffu, b0 ? v[1] : v[3], 1);
    const double a = (b0 ? v[2] : v[0]) + r0;
    const double b = (b0 ? v[3] : v[1]) + r1;
    const double r2 = __shfl_xor_sync(0xffffffffu, b1 ? a : b, 2);
    double t = (b1 ? b : a) + r2;
#pragma unroll
    for (int m = 4; m < 32; m <<= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
    // lane l now holds the frame's total of stream ((l & 1) << 1) | ((l >> 1) & 1); lanes 4 j .. 4 j + 3 keep frame j of 8
    if ((lane >> 2) == (k & 7)) *mine = t;
}

template <>
__device__ __forceinline__ void reduce_and_keep<8>(double (&v)[8], int lane, int k, double* mine) {
    // distances 1, 2, 4 halve the eight streams' values over the lanes (bit 0 picks streams 4..7, bit 1 the upper pair of
    // the four left, bit 2 the upper one of the two left), distances 8 and 16 finish the sum: the tree of reduce_and_keep<1>
    const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4;
    double a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = (b0 ? v[4 + i] : v[i]) + __shfl_xor_sync(0xffffffffu, b0 ? v[i] : v[4 + i], 1);
    double b[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) b[i] = (b1 ? a[2 + i] : a[i]) + __shfl_xor_sync(0xffffffffu, b1 ? a[i] : a[2 + i], 2);
    double t = (b2 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, b2 ? b[0] : b[1], 4);
#pragma unroll
    for (int m = 8; m < 32; m <<= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
    // lane l holds the frame's total of stream 4 (l & 1) + 2 ((l >> 1) & 1) + ((l >> 2) & 1); lanes 8 j .. 8 j + 7 keep frame j of 4
    if ((lane >> 3) == (k & 3)) *mine = t;
}

// The phase table is mirror symmetric -- row r, tap p equals row 254 - r, tap sinc_len - 1 - p to one ulp (row 255 mirrors into
// itself) -- so the kernel reads rows 128 .. 254 as the reversed rows 126 .. 0: the rows it touches (128 KB at 128 taps) stay
// in L1 next to the streamed input, where the whole table (256 KB) did not (L1 hit rate 65 %).
template <int T, int S, int DIR>
__device__ __forceinline__ void accumulate_rows(const double* __restrict__ r0, const double* __restrict__ xx, int span, double (&acc)[S][4]) {
    constexpr int SL = 32 * T;
#pragma unroll
    for (int j = 0; j < T; ++j) {
        double tv[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) tv[d] = __ldg(r0 + DIR * (d * SL + 32 * j));
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double xv = xx[s * span + 32 * j];
#pragma unroll
            for (int d = 0; d < 4; ++d) acc[s][d] = fma(xv, tv[d], acc[s][d]);
        }
    }
}

template <int T, int S>
__global__ void __launch_bounds__(kResampleFramesPerBlock) k_resample(const double* __restrict__ in, size_t in_stride, long long n_in,
                                                                      double* __restrict__ out, size_t out_stride, long long n_frames,
                                                                      const ResampleFrame* __restrict__ frames,
                                                                      const double* __restrict__ table, int n_streams) {
    constexpr int SL = 32 * T;
    extern __shared__ double xs[];
    const long long f0 = static_cast<long long>(blockIdx.x) * kResampleFramesPerBlock;
    const int nf = static_cast<int>(min(static_cast<long long>(kResampleFramesPerBlock), n_frames - f0));
    const int s_base = blockIdx.y * S;
    const int lo = frames[f0].base - 1;
    const int span = frames[f0 + nf - 1].base + SL + 2 - lo;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const bool live = s_base + s < n_streams;
        const double* x = in + static_cast<size_t>(live ? s_base + s : s_base) * in_stride;
        for (int i = threadIdx.x; i < span; i += kResampleFramesPerBlock) {
            const long long p = static_cast<long long>(lo) + i;
            xs[s * span + i] = (live && p >= 0 && p < n_in) ? x[p] : 0.0;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double mine = 0.0;
    for (int k = 0; k < 32; ++k) {
        const int fi = warp * 32 + k;
        if (fi >= nf) break;
        const ResampleFrame fr = frames[f0 + fi];
        const double t = fr.frac, t2 = t * t, t3 = t2 * t;
        // interp_cubic (rubato): a0 = y1, a1 = -y0/3 - y1/2 + y2 - y3/6, a2 = (y0 + y2)/2 - y1, a3 = (y1 - y2)/2 + (y3 - y0)/6
        const double w[4] = {-(1.0 / 3.0) * t + 0.5 * t2 - (1.0 / 6.0) * t3, 1.0 - 0.5 * t - t2 + 0.5 * t3, t + 0.5 * t2 - 0.5 * t3,
                             -(1.0 / 6.0) * t + (1.0 / 6.0) * t3};
        double acc[S][4];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int d = 0; d < 4; ++d) acc[s][d] = 0.0;
        const double* xx = xs + (fr.base - lo) + lane;
        if (fr.sub >= 1 && fr.sub + 2 <= 127) {  // phases sub - 1 .. sub + 2: four consecutive rows over one input window
            accumulate_rows<T, S, 1>(table + static_cast<size_t>(fr.sub - 1) * SL + lane, xx, span, acc);
        } else if (fr.sub - 1 >= 128 && fr.sub + 2 <= 254) {  // the same through the mirror: rows 254 - r, taps reversed
            accumulate_rows<T, S, -1>(table + static_cast<size_t>(254 - (fr.sub - 1)) * SL + (SL - 1 - lane), xx, span, acc);
        } else {  // a phase wraps into the neighbouring sample (3 of 256 positions) or the four rows straddle the mirror
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                int sd = fr.sub + d - 1;
                const int carry = sd < 0 ? -1 : (sd >= kResamplePhases ? 1 : 0);
                sd -= carry * kResamplePhases;
                const bool mirrored = sd >= 128 && sd <= 254;
                const double* row = table + static_cast<size_t>(mirrored ? 254 - sd : sd) * SL + (mirrored ? SL - 1 - lane : lane);
                const int step = mirrored ? -32 : 32;
#pragma unroll
                for (int j = 0; j < T; ++j) {
                    const double tv = __ldg(row + step * j);
#pragma unroll
                    for (int s = 0; s < S; ++s) acc[s][d] = fma(xx[s * span + carry + 32 * j], tv, acc[s][d]);
                }
            }
        }
        double part[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            part[s] = 0.0;
#pragma unroll
            for (int d = 0; d < 4; ++d) part[s] = fma(w[d], acc[s][d], part[s]);
        }
        reduce_and_keep<S>(part, lane, k, &mine);
        if (S == 8 && ((k & 3) == 3 || fi == nf - 1)) {  // four frames x eight streams sit in the warp
            const int frame = (fi & ~3) + (lane >> 3);
            const int stream = s_base + (((lane & 1) << 2) | (lane & 2) | ((lane >> 2) & 1));
            if (frame < nf && stream < n_streams) out[static_cast<size_t>(stream) * out_stride + f0 + frame] = mine;
        }
        if (S == 4 && ((k & 7) == 7 || fi == nf - 1)) {  // eight frames x four streams sit in the warp: 64-byte runs per stream
            const int frame = (fi & ~7) + (lane >> 2);
            const int stream = s_base + (((lane & 1) << 1) | ((lane >> 1) & 1));
            if (frame < nf && stream < n_streams) out[static_cast<size_t>(stream) * out_stride + f0 + frame] = mine;
        }
    }
    if (S == 1 && warp * 32 + lane < nf) out[static_cast<size_t>(s_base) * out_stride + f0 + warp * 32 + lane] = mine;
}

// any sinc_len (a multiple of 32), one stream per CTA: the fallback of the shapes without an unrolled instance
__global__ void __launch_bounds__(kResampleFramesPerBlock) k_resample_any(const double* __restrict__ in, size_t in_stride, long long n_in,
                                                                          double* __restrict__ out, size_t out_stride, long long n_frames,
                                                                          const ResampleFrame* __restrict__ frames,
                                                                          const double* __restrict__ table, int sinc_len) {
    extern __shared__ double xs[];
    const long long f0 = static_cast<long long>(blockIdx.x) * kResampleFramesPerBlock;
    const int nf = static_cast<int>(min(static_cast<long long>(kResampleFramesPerBlock), n_frames - f0));
    const double* x = in + static_cast<size_t>(blockIdx.y) * in_stride;
    const int lo = frames[f0].base - 1;
    const int span = frames[f0 + nf - 1].base + sinc_len + 2 - lo;
    for (int i = threadIdx.x; i < span; i += kResampleFramesPerBlock) {
        const long long p = static_cast<long long>(lo) + i;
        xs[i] = (p >= 0 && p < n_in) ? x[p] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double mine = 0.0;
    for (int k = 0; k < 32; ++k) {
        const int fi = warp * 32 + k;
        if (fi >= nf) break;
        const ResampleFrame fr = frames[f0 + fi];
        const double t = fr.frac, t2 = t * t, t3 = t2 * t;
        const double w[4] = {-(1.0 / 3.0) * t + 0.5 * t2 - (1.0 / 6.0) * t3, 1.0 - 0.5 * t - t2 + 0.5 * t3, t + 0.5 * t2 - 0.5 * t3,
                             -(1.0 / 6.0) * t + (1.0 / 6.0) * t3};
        double part[1] = {0.0};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            int sd = fr.sub + d - 1;
            const int carry = sd < 0 ? -1 : (sd >= kResamplePhases ? 1 : 0);
            sd -= carry * kResamplePhases;
            const double* row = table + static_cast<size_t>(sd) * sinc_len;
            const double* xx = xs + (fr.base - lo + carry);
            double acc = 0.0;
            for (int p = lane; p < sinc_len; p += 32) acc = fma(xx[p], __ldg(row + p), acc);
            part[0] = fma(w[d], acc, part[0]);
        }
        reduce_and_keep<1>(part, lane, k, &mine);
    }
    if (warp * 32 + lane < nf) out[static_cast<size_t>(blockIdx.y) * out_stride + f0 + warp * 32 + lane] = mine;
}

namespace {
template <int T, int S>
cudaError_t launch_instance(const double* d_in, size_t in_stride, size_t n_in, double* d_out, size_t out_stride, size_t n_frames, int first,
                            int n_streams, const ResampleFrame* d_frames, const double* d_table, int max_span, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(max_span) * S * sizeof(double);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        const cudaError_t err = cudaFuncSetAttribute(k_resample<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (err != cudaSuccess) return err;
    }
    const unsigned groups = static_cast<unsigned>((n_frames + kResampleFramesPerBlock - 1) / kResampleFramesPerBlock);
    const int cta_rows = (n_streams + S - 1) / S;
    for (int r0 = 0; r0 < cta_rows; r0 += 65535) {  // gridDim.y limit
        const int rows = std::min(65535, cta_rows - r0);
        const int s0 = first + r0 * S;
        k_resample<T, S><<<dim3(groups, rows), kResampleFramesPerBlock, smem, stream>>>(
            d_in + static_cast<size_t>(s0) * in_stride, in_stride, static_cast<long long>(n_in), d_out + static_cast<size_t>(s0) * out_stride,
            out_stride, static_cast<long long>(n_frames), d_frames, d_table, n_streams - r0 * S);
    }
    return cudaGetLastError();
}

template <int T>
cudaError_t launch_taps(const double* d_in, size_t in_stride, size_t n_in, double* d_out, size_t out_stride, size_t n_frames, int n_streams,
                        const ResampleFrame* d_frames, const double* d_table, int max_span, cudaStream_t stream) {
    // groups of eight, then four streams share the phase-table loads; what is left over renders one stream per CTA
    // (AFSIM_RESAMPLE_GROUP = 4 / 1 caps the group size: measurement knob, read per call)
    const char* env = std::getenv("AFSIM_RESAMPLE_GROUP");
    const int max_group = env ? std::atoi(env) : 8;
    int done = 0;
    if (max_group >= 8 && n_streams - done >= 8) {
        const int n = (n_streams - done) / 8 * 8;
        const cudaError_t err = launch_instance<T, 8>(d_in, in_stride, n_in, d_out, out_stride, n_frames, done, n, d_frames, d_table, max_span, stream);
        if (err != cudaSuccess) return err;
        done += n;
    }
    if (max_group >= 4 && n_streams - done >= 4) {
        const int n = (n_streams - done) / 4 * 4;
        const cudaError_t err = launch_instance<T, 4>(d_in, in_stride, n_in, d_out, out_stride, n_frames, done, n, d_frames, d_table, max_span, stream);
        if (err != cudaSuccess) return err;
        done += n;
    }
    if (n_streams > done)
        return launch_instance<T, 1>(d_in, in_stride, n_in, d_out, out_stride, n_frames, done, n_streams - done, d_frames, d_table, max_span,
                                     stream);
    return cudaSuccess;
}
}  // namespace

cudaError_t launch_resample(const double* d_in, size_t in_stride, size_t n_in, double* d_out, size_t out_stride, size_t n_frames,
                            int n_streams, const ResampleFrame* d_frames, const double* d_table, int sinc_len, int max_span,
                            cudaStream_t stream) {
    if (n_frames == 0 || n_streams == 0) return cudaSuccess;
    if (sinc_len == 128) return launch_taps<4>(d_in, in_stride, n_in, d_out, out_stride, n_frames, n_streams, d_frames, d_table, max_span, stream);
    if (sinc_len == 256) return launch_taps<8>(d_in, in_stride, n_in, d_out, out_stride, n_frames, n_streams, d_frames, d_table, max_span, stream);
    const size_t smem = static_cast<size_t>(max_span) * sizeof(double);
    if (smem > 48 * 1024) {
        const cudaError_t err = cudaFuncSetAttribute(k_resample_any, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (err != cudaSuccess) return err;
    }
    const unsigned groups = static_cast<unsigned>((n_frames + kResampleFramesPerBlock - 1) / kResampleFramesPerBlock);
    for (int s0 = 0; s0 < n_streams; s0 += 65535) {  // gridDim.y limit
        const int ns = std::min(65535, n_streams - s0);
        k_resample_any<<<dim3(groups, ns), kResampleFramesPerBlock, smem, stream>>>(d_in + static_cast<size_t>(s0) * in_stride, in_stride,
                                                                                    static_cast<long long>(n_in),
                                                                                    d_out + static_cast<size_t>(s0) * out_stride, out_stride,
                                                                                    static_cast<long long>(n_frames), d_frames, d_table, sinc_len);
    }
    return cudaGetLastError();
}

}  // namespace afsim
