// afsim_render.h -- per-stream bodies of the stage kernels and the per-stream finalize reduction.
//
// body_*(args, chunk, s, ...) is what ONE CUDA thread does for stream s over one chunk; the
// __global__ wrappers in afsim_kernels.cu only compute `s` and call it.  Keeping the bodies host /
// device portable lets tests/hostsim run the identical chunked schedule on the CPU (test harness
// only -- libafsim.so contains no CPU path).
//
// Reference citations are relative to rust-core/src/.
#pragma once
#include "../../include/afsim.h"
#include "afsim_cleanup.h"
#include "afsim_deesser.h"
#include "afsim_split.h"

namespace afsim {

struct ChunkArgs {
    int n0;    // first sample of the chunk
    int len;   // samples in the chunk
    int row0;  // ring row of sample n0
};

typedef float FirTable[4][32];

AF_HD const CandidateParams& stream_params(const BatchArgs& a, int s) { return a.params[a.cand[s]]; }

// ---- input -------------------------------------------------------------------------------------------------
// Adaptive cleanup (AF_INPUT_CLEANUP_*): per 480-sample block, analyse the raw block, then DC block + notches
// + adaptive high-pass (processor/tests.rs:500-549).  Chunks are multiples of 480 in this mode.
AF_HD void body_input_cleanup(const BatchArgs& a, const ChunkArgs& ck, int s) {
    const size_t stride = (size_t)a.stride;
    const CleanupConst& k = *a.cleanup;
    const bool gentle = a.input_stage == AF_INPUT_CLEANUP_GENTLE;
    InputStage st;
    CleanupStage cl;
    if (ck.n0 == 0) {
        st.init();
        cl.init(k);
    } else {
        StateIO<false> io{a.st_input + s, stride};
        st.sync(io);
        cl.sync(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    const Col out{a.buf_a + (size_t)ck.row0 * stride + s, stride};
    const float* src = a.signals + a.src_off[s];
    float* rows_in = a.rows + s;
    for (int b0 = 0; b0 < ck.len; b0 += kInputBlock) {
        const int blen = ck.len - b0 < kInputBlock ? ck.len - b0 : kInputBlock;
        for (int t = 0; t < blen; ++t) {  // analyze_input on the raw (sanitised) block
            float v = src[ck.n0 + b0 + t];
            if (!af_finite(v)) v = 0.0f;
            out.set(b0 + t, v);
            cl.analyze(v, k, gentle);
        }
        bool hum_detected;
        cl.begin_block(k, gentle, &hum_detected);
        for (int t = 0; t < blen; ++t) {
            const int n = ck.n0 + b0 + t;
            const float in = out.get(b0 + t);
            const float dc = in - st.x1 + 0.995f * st.y1;  // routing.rs:832-836 (no fixed high-pass in this mode)
            st.x1 = in;
            st.y1 = dc;
            float v = cl.process(dc, k);
            if (!af_finite(v)) v = 0.0f;
            const double sq = (double)v * (double)v;
            st.sum_in += sq;
            st.blk_in += sq;
            st.peak_in = fmaxf(st.peak_in, fabsf(v));
            out.set(b0 + t, v);
            if (clk.at_end(n)) {
                const float rms = (float)sqrt(st.blk_in / (double)clk.block_len(n));
                rows_in[(size_t)clk.blk * stride] = lin_to_db_f32(rms);
                st.blk_in = 0.0;
                clk.advance();
            }
        }
    }
    if (ck.n0 + ck.len >= a.n_samples) {
        a.accum[s].sum_in = st.sum_in;
        a.accum[s].peak_in = st.peak_in;
    } else {
        StateIO<true> io{a.st_input + s, stride};
        st.sync(io);
        cl.sync(io);
    }
}

AF_HD bool input_uses_cleanup(const BatchArgs& a) {
    return a.input_stage == AF_INPUT_CLEANUP_GENTLE || a.input_stage == AF_INPUT_CLEANUP_STRONG;
}

AF_HD void body_input(const BatchArgs& a, const ChunkArgs& ck, int s) {
    const size_t stride = (size_t)a.stride;
    InputStage st;
    if (ck.n0 == 0) {
        st.init();
    } else {
        StateIO<false> io{a.st_input + s, stride};
        st.sync(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    const Col out{a.buf_a + (size_t)ck.row0 * stride + s, stride};
    const float* src = a.signals + a.src_off[s];
    float* rows_in = a.rows + s;  // row.0
    if (a.input_stage == AF_INPUT_DC_HP80) {
        const Bq hp = bq_from(stream_params(a, s).in_hp);
        st.run<true>(src, out, ck.n0, ck.len, hp, clk, rows_in, stride);
    } else {
        Bq hp;
        hp.b0 = 1.0;
        hp.b1 = hp.b2 = hp.a1 = hp.a2 = 0.0;
        st.run<false>(src, out, ck.n0, ck.len, hp, clk, rows_in, stride);
    }
    if (ck.n0 + ck.len >= a.n_samples) {
        a.accum[s].sum_in = st.sum_in;
        a.accum[s].peak_in = st.peak_in;
    } else {
        StateIO<true> io{a.st_input + s, stride};
        st.sync(io);
    }
}

// ---- shared input stage: copy the distinct passage's input stage output, rows and statistics to stream s -----------
constexpr int kFanoutGroup = 32;  // samples per thread
AF_HD void body_input_fanout(const BatchArgs& a, const ChunkArgs& ck, int s, int g) {
    const size_t stride = (size_t)a.stride, ustride = (size_t)a.in_stride;
    const uint32_t u = a.in_unique[s];
    const int t0 = g * kFanoutGroup;
    const int valid = ck.len - t0 < kFanoutGroup ? ck.len - t0 : kFanoutGroup;
    if (!a.in_det && !a.in_de[0]) {  // with a shared compressor / de-esser front the streams read the shared signal in place
        const float* src = a.in_src + (size_t)(ck.row0 + t0) * ustride + u;
        float* dst = a.buf_a + (size_t)(ck.row0 + t0) * stride + s;
        for (int t = 0; t < valid; ++t) dst[(size_t)t * stride] = src[(size_t)t * ustride];
    }
    if (g != 0) return;
    const int end = ck.n0 + ck.len;  // analysis blocks that end inside this chunk
    for (int b = ck.n0 / a.block_samples; b < a.n_rows; ++b) {
        const int block_end = (b + 1) * a.block_samples < a.n_samples ? (b + 1) * a.block_samples : a.n_samples;
        if (block_end > end) break;
        if (block_end > ck.n0) a.rows[(size_t)b * stride + s] = a.in_rows[(size_t)b * ustride + u];
    }
    if (end >= a.n_samples) {
        a.accum[s].sum_in = a.in_accum[u].sum_in;
        a.accum[s].peak_in = a.in_accum[u].peak_in;
    }
}

// ---- EQ slice: sections [first, first + K) -------------------------------------------------------------------
template <int K>
AF_HD void body_eq(const BatchArgs& a, const ChunkArgs& ck, int s, int first) {
    const size_t stride = (size_t)a.stride;
    const CandidateParams& p = stream_params(a, s);
    if ((int)p.n_sections <= first) return;  // nothing of this slice in the stream's cascade: samples pass through
    EqStage<K> st;
    st.init(p, first);
    double* table = a.st_eq + (size_t)(kStateEqPerSection * first) * stride + s;
    if (ck.n0 != 0) {
        StateIO<false> io{table, stride};
        st.sync(io);
    }
    const Col io{a.buf_a + (size_t)ck.row0 * stride + s, stride};
    int t_begin = 0;
    if (ck.n0 == 0 && (p.flags & LF_EQ_FADE)) t_begin = st.run_fade_head(io, ck.len, a.fade_samples, first, a.eq_default);
    st.run(io, t_begin, ck.len);
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> sio{table, stride};
        st.sync(sio);
    }
}

// ---- compressor -----------------------------------------------------------------------------------------------
AF_HD void body_compressor(const BatchArgs& a, const ChunkArgs& ck, int s) {
    const size_t stride = (size_t)a.stride;
    CompressorStage st;
    st.init(stream_params(a, s));
    if (ck.n0 != 0) {
        StateIO<false> io{a.st_comp + s, stride};
        st.sync(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    const Col io{a.buf_a + (size_t)ck.row0 * stride + s, stride};
    st.run(io, ck.n0, ck.len, clk, a.rows + (size_t)2 * a.n_rows * stride + s, stride);
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> sio{a.st_comp + s, stride};
        st.sync(sio);
    }
}

// Compressor of a stream whose front (sidechain, detector weight, instantaneous peak) and input (the EQ output) come
// from the shared render of its (passage, EQ) pair: shared rings -> buf_a.
AF_HD void body_compressor_shared(const BatchArgs& a, const ChunkArgs& ck, int s) {
    const size_t stride = (size_t)a.stride, ustride = (size_t)a.in_stride;
    CompressorStage st;
    st.init(stream_params(a, s));
    if (ck.n0 != 0) {
        StateIO<false> io{a.st_comp + s, stride};
        st.sync(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    const size_t o = (size_t)ck.row0 * ustride + a.in_unique[s];
    const Col out{a.buf_a + (size_t)ck.row0 * stride + s, stride};
    st.run_shared(a.in_src + o, a.in_det + o, a.in_wdb + o, a.in_ipk + o, ustride, out, ck.n0, ck.len, clk,
                  a.rows + (size_t)2 * a.n_rows * stride + s, stride);
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> sio{a.st_comp + s, stride};
        st.sync(sio);
    }
}

// ---- sample limiter: buf_a -> buf_b -----------------------------------------------------------------------------
AF_HD void body_limiter(const BatchArgs& a, const ChunkArgs& ck, int s) {
    const size_t stride = (size_t)a.stride;
    const CandidateParams& p = stream_params(a, s);
    LimiterStage st;
    if (ck.n0 == 0) {
        st.init();
    } else {
        StateIO<false> io{a.st_lim + s, stride};
        st.sync(io);
    }
    st.run(a.buf_a + s, a.buf_b + s, stride, a.ring_rows, ck.row0, ck.n0, ck.len, a.lookahead, p.l_ceil, p.l_release,
           a.lim_sfx + s);
    if (ck.n0 + ck.len >= a.n_samples) {
        a.accum[s].limiter_gr_db = st.peak_reduction_db();
    } else {
        StateIO<true> sio{a.st_lim + s, stride};
        st.sync(sio);
    }
}

// ---- true-peak limiter + detector + output statistics (LIMITER = false: detector + statistics only) -------------
template <bool LIMITER>
AF_HD void body_output(const BatchArgs& a, const ChunkArgs& ck, int s, const FirTable& fir) {
    const size_t stride = (size_t)a.stride;
    const CandidateParams& p = stream_params(a, s);
    TpLimiterStage tp;
    OutputStage out;
    if (ck.n0 == 0) {
        tp.init();
        out.init();
    } else {
        StateIO<false> io{a.st_tp + s, stride};
        if (LIMITER) tp.sync(io);
        out.sync(io);
    }
    BlockClock clk_tp, clk_out;
    clk_tp.init(a.block_samples, a.n_samples, ck.n0);
    clk_out = clk_tp;
    const float ceil_lin = p.tp_ceil, rel = p.tp_release, one_m_rel = 1.0f - p.tp_release;
    const float* in = (LIMITER ? a.buf_b : a.buf_a) + (size_t)ck.row0 * stride + s;
    float* audio = a.audio ? a.audio + a.audio_off[s] + ck.n0 : nullptr;
    float* rows_out = a.rows + (size_t)1 * a.n_rows * stride + s;
    for (int t0 = 0; t0 < ck.len; t0 += kFirChunk) {
        const int valid = ck.len - t0 < kFirChunk ? ck.len - t0 : kFirChunk;
        float x[kFirChunk], y[kFirChunk];
#pragma unroll
        for (int j = 0; j < kFirChunk; ++j) x[j] = j < valid ? in[(size_t)(t0 + j) * stride] : 0.0f;
        if (LIMITER) {
            tp.group(x, y, valid, ck.n0 + t0, clk_tp, fir, ceil_lin, rel, one_m_rel);
        } else {
#pragma unroll
            for (int j = 0; j < kFirChunk; ++j) y[j] = x[j];
        }
        out.group(y, valid, ck.n0 + t0, clk_out, fir, rows_out, stride);
        if (audio) {
#pragma unroll
            for (int j = 0; j < kFirChunk; ++j)
                if (j < valid) audio[t0 + j] = y[j];
        }
    }
    if (ck.n0 + ck.len >= a.n_samples) {
        StreamAccum& acc = a.accum[s];
        acc.sum_out = out.sum_out;
        acc.peak_out = out.peak_out;
        acc.peak_out_tp = out.peak_tp;
        acc.non_finite = out.non_finite ? 1u : 0u;
        if (LIMITER) {
            acc.peak_pre_tp = tp.peak_pre;
            acc.tp_gr_db = tp.peak_reduction_db();
            acc.events = tp.events;
        }
    } else {
        StateIO<true> sio{a.st_tp + s, stride};
        if (LIMITER) tp.sync(sio);
        out.sync(sio);
    }
}

// ---- simulate_eq_v2: true-peak detector over the (sanitised) input held in buf_a, before the EQ ----------------
AF_HD void body_input_true_peak(const BatchArgs& a, const ChunkArgs& ck, int s, const FirTable& fir) {
    const size_t stride = (size_t)a.stride;
    OutputStage det;
    double* table = a.st_tp + (size_t)40 * stride + s;  // second half of the true-peak state table
    if (ck.n0 == 0) {
        det.init();
    } else {
        StateIO<false> io{table, stride};
        det.sync(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    const float* in = a.buf_a + (size_t)ck.row0 * stride + s;
    for (int t0 = 0; t0 < ck.len; t0 += kFirChunk) {
        const int valid = ck.len - t0 < kFirChunk ? ck.len - t0 : kFirChunk;
        float x[kFirChunk];
#pragma unroll
        for (int j = 0; j < kFirChunk; ++j) x[j] = j < valid ? in[(size_t)(t0 + j) * stride] : 0.0f;
        det.group(x, valid, ck.n0 + t0, clk, fir, nullptr, stride);
    }
    if (ck.n0 + ck.len >= a.n_samples) {
        a.accum[s].peak_in_tp = det.peak_tp;
    } else {
        StateIO<true> sio{table, stride};
        det.sync(sio);
    }
}

// =====================================================================================================================
// split (R/M) path: see afsim_split.h.  R bodies: one call per stream and chunk; M bodies: one call per
// (stream, group of kGroup samples).
// =====================================================================================================================
AF_HD bool group_span(const ChunkArgs& ck, int g, int* t0, int* valid, int group = kGroup) {
    *t0 = g * group;
    *valid = ck.len - *t0 < group ? ck.len - *t0 : group;
    return *valid > 0;
}
template <typename T>
AF_HD T* col_at(T* ring, const BatchArgs& a, const ChunkArgs& ck, int s, int t0 = 0) {
    return ring + (size_t)(ck.row0 + t0) * (size_t)a.stride + s;
}

AF_HD void atomic_max_nonneg(float* addr, float v) {
    if (!(v == v)) return;  // f32::max ignores NaN
#if defined(__CUDA_ARCH__)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
#else
    if (v > *addr) *addr = v;
#endif
}

AF_HD void body_comp_r1(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {
    const size_t stride = (size_t)a.stride;
    CompSplit st;
    st.init(stream_params(a, s));
    if (ck.n0 != 0) {
        StateIO<false> io{a.st_comp + s, stride};
        st.sync_r1(io);
    }
    st.run_r1(col_at(a.buf_a, a, ck, s), col_at(a.w[0], a, ck, s), col_at(a.w[1], a, ck, s), col_at(a.w[2], a, ck, s),
              col_at(a.w[3], a, ck, s), stride, ck.len, stg);
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> io{a.st_comp + s, stride};
        st.sync_r1(io);
    }
}
AF_HD void body_comp_m2(const BatchArgs& a, const ChunkArgs& ck, int s, int g) {
    int t0, valid;
    if (!group_span(ck, g, &t0, &valid, kCompMapGroup)) return;
    CompSplit st;
    st.init_map(a.map_tab + s, (size_t)a.stride);
    st.map_m2(col_at(a.w[0], a, ck, s, t0), col_at(a.w[1], a, ck, s, t0), col_at(a.w[2], a, ck, s, t0),
              col_at(a.w[3], a, ck, s, t0), (size_t)a.stride, valid);
}
AF_HD void body_comp_r3(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {
    const size_t stride = (size_t)a.stride;
    CompSplit st;
    st.init(stream_params(a, s));
    double* table = a.st_comp + (size_t)5 * stride + s;
    if (ck.n0 != 0) {
        StateIO<false> io{table, stride};
        st.sync_r3(io);
    }
    if (a.in_det) {  // sidechain signal and instantaneous peak of the shared compressor front
        const size_t o = (size_t)ck.row0 * (size_t)a.in_stride + a.in_unique[s];
        st.run_r3(a.in_det + o, a.in_ipk + o, (size_t)a.in_stride, col_at(a.w[2], a, ck, s), col_at(a.w[3], a, ck, s), stride, ck.len, stg);
    } else {
        st.run_r3(col_at(a.w[0], a, ck, s), col_at(a.w[2], a, ck, s), stride, col_at(a.w[2], a, ck, s), col_at(a.w[3], a, ck, s), stride,
                  ck.len, stg);
    }
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> io{table, stride};
        st.sync_r3(io);
    }
}
AF_HD void body_comp_m4(const BatchArgs& a, const ChunkArgs& ck, int s, int g) {
    int t0, valid;
    if (!group_span(ck, g, &t0, &valid, kCompMapGroup)) return;
    CompSplit st;
    st.init_map(a.map_tab + s, (size_t)a.stride);
    double* w1 = col_at(a.w[1], a, ck, s, t0);
    if (a.in_det)
        st.map_m4(a.in_wdb + (size_t)(ck.row0 + t0) * (size_t)a.in_stride + a.in_unique[s], (size_t)a.in_stride, w1,
                  col_at(a.w[2], a, ck, s, t0), col_at(a.w[3], a, ck, s, t0), (size_t)a.stride, valid);
    else
        st.map_m4(w1, (size_t)a.stride, w1, col_at(a.w[2], a, ck, s, t0), col_at(a.w[3], a, ck, s, t0), (size_t)a.stride, valid);
}
AF_HD void body_comp_r5(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {
    const size_t stride = (size_t)a.stride;
    CompSplit st;
    st.init(stream_params(a, s));
    double* table = a.st_comp + (size_t)7 * stride + s;
    if (ck.n0 != 0) {
        StateIO<false> io{table, stride};
        st.sync_r5(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    st.run_r5(col_at(a.w[1], a, ck, s), stride, ck.n0, ck.len, clk, a.rows + (size_t)2 * a.n_rows * stride + s, stg);
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> io{table, stride};
        st.sync_r5(io);
    }
}
AF_HD void body_comp_m6(const BatchArgs& a, const ChunkArgs& ck, int s, int g) {
    int t0, valid;
    if (!group_span(ck, g, &t0, &valid, kCompMapGroup)) return;
    CompSplit st;
    st.init_map(a.map_tab + s, (size_t)a.stride);
    if (a.structure & ST_AUTO_MAKEUP)
        st.map_m6_gain(col_at(a.w[1], a, ck, s, t0), (size_t)a.stride, valid);
    else
    {
        float* x = col_at(a.buf_a, a, ck, s, t0);
        if (a.in_det)  // the compressor's input is the shared EQ output
            st.map_m6(col_at(a.w[1], a, ck, s, t0), a.in_src + (size_t)(ck.row0 + t0) * (size_t)a.in_stride + a.in_unique[s],
                      (size_t)a.in_stride, x, (size_t)a.stride, valid);
        else
            st.map_m6(col_at(a.w[1], a, ck, s, t0), x, (size_t)a.stride, x, (size_t)a.stride, valid);
    }
}
// R7 (auto-makeup batches): apply gain reduction x makeup, run the loudness meter, step the makeup at block ends
AF_HD void body_comp_r7(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {
    const size_t stride = (size_t)a.stride;
    MakeupR st;
    st.init(stream_params(a, s));
    if (ck.n0 != 0) {
        StateIO<false> io{a.st_mk + s, stride};
        st.sync(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    const int64_t voff = a.mk_vad_off ? a.mk_vad_off[s] : -1;
    st.run(col_at(a.buf_a, a, ck, s), col_at(a.w[1], a, ck, s), stride, ck.n0, ck.len, clk, *a.mk_const, a.mk_ring + s,
           a.mk_rows ? a.mk_rows + s : nullptr, a.n_rows, (a.mk_vad && voff >= 0) ? a.mk_vad + voff : nullptr, stg);
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> io{a.st_mk + s, stride};
        st.sync(io);
    }
}

// ---- de-esser: R_a -> M_b -> R_c1 -> M_c2 -> R_c3 (afsim_deesser.h), used by every batch that has the stage ---------------------------
AF_HD void body_de_ra(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {
    const size_t stride = (size_t)a.stride;
    const CandidateParams& p = stream_params(a, s);
    DeEsserDetect st;
    if (ck.n0 == 0) {
        st.init();
    } else {
        StateIO<false> io{a.st_deesser + s, stride};
        st.sync(io);
    }
    const DeConst k{a.de_tab + s, stride};
    st.run(col_at(a.buf_a, a, ck, s), col_at(a.w[0], a, ck, s), col_at(a.w[1], a, ck, s), col_at(a.w[2], a, ck, s),
           col_at(a.w[3], a, ck, s), stride, ck.n0, ck.len, a.fade_samples, k, &p, stg);
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> io{a.st_deesser + s, stride};
        st.sync(io);
    }
}
AF_HD void body_de_mb(const BatchArgs& a, const ChunkArgs& ck, int s, int g) {
    int t0, valid;
    if (!group_span(ck, g, &t0, &valid, kDeMapGroup)) return;
    deesser_levels(col_at(a.w[0], a, ck, s, t0), col_at(a.w[1], a, ck, s, t0), col_at(a.w[2], a, ck, s, t0),
                   col_at(a.w[3], a, ck, s, t0), col_at(a.w[4], a, ck, s, t0), col_at(a.w[5], a, ck, s, t0),
                   col_at(a.w[6], a, ck, s, t0), (size_t)a.stride, valid);
}
AF_HD void body_de_rc(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {  // R_c1
    const size_t stride = (size_t)a.stride;
    DeEsserTargets st;
    st.init(stream_params(a, s));
    double* table = a.st_deesser + (size_t)kStateDeDetect * stride + s;
    if (ck.n0 != 0) {
        StateIO<false> io{table, stride};
        st.sync(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    const DeConst k{a.de_tab + s, stride};
    double* const w[7] = {col_at(a.w[0], a, ck, s), col_at(a.w[1], a, ck, s), col_at(a.w[2], a, ck, s), col_at(a.w[3], a, ck, s),
                          col_at(a.w[4], a, ck, s), col_at(a.w[5], a, ck, s), col_at(a.w[6], a, ck, s)};
    if (a.in_de[0]) {  // levels and confidence targets of the shared front
        const size_t o = (size_t)ck.row0 * (size_t)a.in_stride + a.in_unique[s];
        const double* const in[7] = {a.in_de[0] + o, a.in_de[1] + o, a.in_de[2] + o, a.in_de[3] + o,
                                     a.in_de[4] + o, a.in_de[5] + o, a.in_de[6] + o};
        st.run(in, (size_t)a.in_stride, w, stride, ck.n0, ck.len, k, clk, a.rows + (size_t)3 * a.n_rows * stride + s, stg);
    } else {
        const double* const in[7] = {w[0], w[1], w[2], w[3], w[4], w[5], w[6]};
        st.run(in, stride, w, stride, ck.n0, ck.len, k, clk, a.rows + (size_t)3 * a.n_rows * stride + s, stg);
    }
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> io{table, stride};
        st.sync(io);
    }
}
// R_c1 cut three ways (afsim_deesser.h): R_c1a -> M_c1b -> R_c1c, same state table as body_de_rc
AF_HD void body_de_rc1a(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {
    const size_t stride = (size_t)a.stride;
    DeEsserConfBase st;
    st.init(stream_params(a, s));
    double* table = a.st_deesser + (size_t)kStateDeDetect * stride + s;
    if (ck.n0 != 0) {
        StateIO<false> io{table, stride};
        st.sync(io);
    }
    const DeConst k{a.de_tab + s, stride};
    double* const out[6] = {col_at(a.w[7], a, ck, s),  col_at(a.w[8], a, ck, s),  col_at(a.w[9], a, ck, s),
                            col_at(a.w[10], a, ck, s), col_at(a.w[11], a, ck, s), col_at(a.w[12], a, ck, s)};
    if (a.in_de[0]) {  // levels and confidence targets of the shared front
        const size_t o = (size_t)ck.row0 * (size_t)a.in_stride + a.in_unique[s];
        const double* const in[7] = {a.in_de[0] + o, a.in_de[1] + o, a.in_de[2] + o, a.in_de[3] + o,
                                     a.in_de[4] + o, a.in_de[5] + o, a.in_de[6] + o};
        st.run(in, (size_t)a.in_stride, out, stride, ck.len, k, stg);
    } else {
        const double* const in[7] = {col_at(a.w[0], a, ck, s), col_at(a.w[1], a, ck, s), col_at(a.w[2], a, ck, s), col_at(a.w[3], a, ck, s),
                                     col_at(a.w[4], a, ck, s), col_at(a.w[5], a, ck, s), col_at(a.w[6], a, ck, s)};
        st.run(in, stride, out, stride, ck.len, k, stg);
    }
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> io{table, stride};
        st.sync(io);
    }
}
AF_HD void body_de_mc1b(const BatchArgs& a, const ChunkArgs& ck, int s, int g) {
    int t0, valid;
    if (!group_span(ck, g, &t0, &valid, kDeTargetGroup)) return;
    const size_t stride = (size_t)a.stride;
    const DeConst k{a.de_tab + s, stride};
    const bool auto_mode = (static_cast<uint32_t>(a.map_tab[(size_t)MT_FLAGS * stride + s]) & LF_DE_AUTO) != 0;
    double* const w[6] = {col_at(a.w[7], a, ck, s, t0),  col_at(a.w[8], a, ck, s, t0),  col_at(a.w[9], a, ck, s, t0),
                          col_at(a.w[10], a, ck, s, t0), col_at(a.w[11], a, ck, s, t0), col_at(a.w[12], a, ck, s, t0)};
    if (a.in_de[0]) {
        const size_t o = (size_t)(ck.row0 + t0) * (size_t)a.in_stride + a.in_unique[s];
        const double* const in[4] = {a.in_de[0] + o, a.in_de[1] + o, a.in_de[2] + o, a.in_de[3] + o};
        deesser_targets(in, (size_t)a.in_stride, w, stride, valid, k, auto_mode);
    } else {
        const double* const in[4] = {col_at(a.w[0], a, ck, s, t0), col_at(a.w[1], a, ck, s, t0), col_at(a.w[2], a, ck, s, t0),
                                     col_at(a.w[3], a, ck, s, t0)};
        deesser_targets(in, stride, w, stride, valid, k, auto_mode);
    }
}
AF_HD void body_de_rc1c(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {
    const size_t stride = (size_t)a.stride;
    DeEsserReduction st;
    st.init();
    double* table = a.st_deesser + (size_t)kStateDeDetect * stride + s;
    if (ck.n0 != 0) {
        StateIO<false> io{table, stride};
        st.sync(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    const DeConst k{a.de_tab + s, stride};
    const double* const tg[3] = {col_at(a.w[7], a, ck, s), col_at(a.w[8], a, ck, s), col_at(a.w[9], a, ck, s)};
    double* const gains[3] = {col_at(a.w[4], a, ck, s), col_at(a.w[5], a, ck, s), col_at(a.w[6], a, ck, s)};
    st.run(tg, col_at(a.w[0], a, ck, s), gains, stride, ck.n0, ck.len, k, clk, a.rows + (size_t)3 * a.n_rows * stride + s, stg);
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> io{table, stride};
        st.sync(io);
    }
}
AF_HD void body_de_mc2(const BatchArgs& a, const ChunkArgs& ck, int s, int g) {
    int t0, valid;
    if (!group_span(ck, g, &t0, &valid, kDeRebuildGroup)) return;
    double* w[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) w[i] = col_at(a.w[i], a, ck, s, t0);
    const DeConst k{a.de_tab + s, (size_t)a.stride};
    deesser_rebuild(w, (size_t)a.stride, valid, k);
}
AF_HD void body_de_rc3(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {
    const size_t stride = (size_t)a.stride;
    const CandidateParams& p = stream_params(a, s);
    DeEsserFilter st;
    st.init(p);
    double* table = a.st_deesser + (size_t)(kStateDeDetect + kStateDeTargets) * stride + s;
    if (ck.n0 != 0) {
        StateIO<false> io{table, stride};
        st.sync(io);
    }
    double* w[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) w[i] = col_at(a.w[i], a, ck, s);
    float* x = col_at(a.buf_a, a, ck, s);
    if (a.in_de[0])  // the de-esser is the first stage: its input is the shared input stage output of the passage
        st.run(a.in_src + (size_t)ck.row0 * (size_t)a.in_stride + a.in_unique[s], (size_t)a.in_stride, x, w, stride, ck.n0, ck.len,
               a.fade_samples, &p, stg);
    else
        st.run(x, stride, x, w, stride, ck.n0, ck.len, a.fade_samples, &p, stg);
    if (ck.n0 + ck.len < a.n_samples) {
        StateIO<true> io{table, stride};
        st.sync(io);
    }
}

AF_HD void body_lim_m(const BatchArgs& a, const ChunkArgs& ck, int s, int g) {
    int t0, valid;
    if (!group_span(ck, g, &t0, &valid, kLimGroup)) return;
    limiter_targets(a.buf_a + s, (size_t)a.stride, a.ring_rows, ck.row0, t0, ck.n0 + t0, valid, a.lookahead,
                    a.map_tab[(size_t)MT_L_CEIL * a.stride + s], col_at(a.w[0], a, ck, s, t0));
}
AF_HD void body_lim_r(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {
    const size_t stride = (size_t)a.stride;
    const CandidateParams& p = stream_params(a, s);
    LimiterR st;
    if (ck.n0 == 0) {
        st.init();
    } else {
        StateIO<false> io{a.st_lim + s, stride};
        st.sync(io);
    }
    st.run(col_at(a.w[0], a, ck, s), a.buf_a + s, col_at(a.buf_b, a, ck, s), stride, a.ring_rows, ck.row0, ck.n0, ck.len,
           a.lookahead, p.l_ceil, p.l_release, stg);
    if (ck.n0 + ck.len >= a.n_samples) {
        a.accum[s].limiter_gr_db = st.peak_reduction_db();
    } else {
        StateIO<true> io{a.st_lim + s, stride};
        st.sync(io);
    }
}


// input true peaks of the true-peak limiter: buf_b -> buf_p
AF_HD void body_tp_fir_in(const BatchArgs& a, const ChunkArgs& ck, int s, int g, const FirTable& fir) {
    int t0, valid;
    if (!group_span(ck, g, &t0, &valid)) return;
    float pk[kFirChunk];
    fir_group_peaks(a.buf_b + s, (size_t)a.stride, a.ring_rows, ck.row0, t0, ck.n0 + t0, valid, fir, pk);
    // feed-forward part of dsp/true_peak.rs:349-354: the target gain of every sample, and the running
    // maximum of the input true peak (order independent -> atomic)
    const float ceil_lin = (float)a.map_tab[(size_t)MT_TP_CEIL * a.stride + s];
    float tgt[kFirChunk];
    float m = 0.0f;
#pragma unroll
    for (int j = 0; j < kFirChunk; ++j) {
        tgt[j] = pk[j] > ceil_lin ? clampf((ceil_lin * 0.999f) / pk[j], 0.0f, 1.0f) : 1.0f;
        if (j < valid) m = fmaxf(m, pk[j]);
    }
    store_tile(col_at(a.buf_p, a, ck, s, t0), (size_t)a.stride, valid, tgt);
    atomic_max_nonneg(&a.accum[s].peak_pre_tp, m);
}
AF_HD void body_tp_r(const BatchArgs& a, const ChunkArgs& ck, int s, Staging stg) {
    const size_t stride = (size_t)a.stride;
    const CandidateParams& p = stream_params(a, s);
    TpR st;
    if (ck.n0 == 0) {
        st.init();
    } else {
        StateIO<false> io{a.st_tp + s, stride};
        st.sync(io);
    }
    BlockClock clk;
    clk.init(a.block_samples, a.n_samples, ck.n0);
    float* audio = a.audio ? a.audio + a.audio_off[s] + ck.n0 : nullptr;
    st.run(col_at(a.buf_p, a, ck, s), a.buf_b + s, col_at(a.buf_c, a, ck, s), audio, stride, a.ring_rows, ck.row0, ck.n0,
           ck.len, p.tp_ceil, p.tp_release, clk, a.rows + (size_t)1 * a.n_rows * stride + s, stg);
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
}
// detector over the true-peak limiter's output: buf_c -> accum.peak_out_tp (order-independent maximum)
AF_HD void body_tp_fir_out(const BatchArgs& a, const ChunkArgs& ck, int s, int g, const FirTable& fir) {
    int t0, valid;
    if (!group_span(ck, g, &t0, &valid)) return;
    float pk[kFirChunk];
    fir_group_peaks(a.buf_c + s, (size_t)a.stride, a.ring_rows, ck.row0, t0, ck.n0 + t0, valid, fir, pk);
    float m = 0.0f;
#pragma unroll
    for (int j = 0; j < kFirChunk; ++j)
        if (j < valid) m = fmaxf(m, pk[j]);
    atomic_max_nonneg(&a.accum[s].peak_out_tp, m);
}

// ---- de-esser constant table: CandidateParams::de -> [DE_FIELDS][S_pad] -------------------------------------------
AF_HD void body_expand_deesser(const BatchArgs& a, int s) {
    const CandidateParams& p = stream_params(a, s);
    double* dst = const_cast<double*>(a.de_tab) + s;
    for (int f = 0; f < DE_FIELDS; ++f) dst[(size_t)f * a.stride] = p.de[f];
}

// =====================================================================================================================
// finalize: per-block rows -> the result dict (audio/processor/python_api.rs:578-713)
// =====================================================================================================================
// Cooperative group of `n` workers (a CUDA thread block; a single worker on the host).
struct Coop {
    int tid, n;
    AF_HD void sync() const {
#if defined(__CUDA_ARCH__)
        __syncthreads();
#endif
    }
};

AF_HD int total_key(float v) {  // f32::total_cmp order as a signed int
    int i;
#if defined(__CUDA_ARCH__)
    i = __float_as_int(v);
#else
    __builtin_memcpy(&i, &v, 4);
#endif
    return i ^ (int)(((unsigned)(i >> 31)) >> 1);
}
AF_HD float key_max_float() {
    const int i = 0x7fffffff;
    float v;
#if defined(__CUDA_ARCH__)
    v = __int_as_float(i);
#else
    __builtin_memcpy(&v, &i, 4);
#endif
    return v;
}

// Bitonic sort of n floats by total_cmp (python_api.rs:62); buf has room for the padded power of two.
AF_HD void coop_sort(const Coop& co, float* buf, int n) {
    int n_pad = 2;
    while (n_pad < n) n_pad <<= 1;
    for (int i = co.tid + n; i < n_pad; i += co.n) buf[i] = key_max_float();
    co.sync();
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = co.tid; i < n_pad; i += co.n) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float x = buf[i], y = buf[ixj];
                    const bool up = (i & k) == 0;
                    const bool gt = total_key(x) > total_key(y);
                    if (gt == up) {
                        buf[i] = y;
                        buf[ixj] = x;
                    }
                }
            }
            co.sync();
        }
    }
}

// percentile_f32 (python_api.rs:58-72) of an already sorted buffer.
AF_HD float sorted_percentile(const float* buf, int n, float p) {
    if (n == 0) return 0.0f;
    const float position = (float)(n - 1) * clampf(p, 0.0f, 1.0f);
    const int lower = (int)floorf(position);
    const int upper = (int)ceilf(position);
    if (lower == upper) return buf[lower];
    const float fraction = position - (float)lower;
    return buf[lower] + fraction * (buf[upper] - buf[lower]);
}

// Workspace of one stream: the four row arrays (n_rows each) + two sort buffers (n_pad each) + 32 scalars.
AF_HD size_t finalize_workspace_floats(int n_rows, int n_pad) { return (size_t)4 * n_rows + (size_t)2 * n_pad + 32; }

enum CompactMode { CM_ACTIVE_GR = 0, CM_ALL_GR, CM_ACTIVE_GAIN, CM_SILENCE_DELTA, CM_SILENCE_GAIN };

// Worker 0 walks the rows (a few thousand at most) and writes the kept values to dst; returns the count
// to every worker through sh[31].
AF_HD int coop_compact(const Coop& co, float* dst, int n_rows, const float* r_in, const float* r_val, const float* r_out,
                       int mode, float thr, float* sh) {
    co.sync();
    if (co.tid == 0) {
        int c = 0;
        for (int i = 0; i < n_rows; ++i) {
            const float in_db = r_in[i];
            bool keep;
            float v;
            switch (mode) {
                case CM_ACTIVE_GR: keep = in_db >= thr; v = fmaxf(r_val[i], 0.0f); break;
                case CM_ALL_GR: keep = true; v = fmaxf(r_val[i], 0.0f); break;
                case CM_ACTIVE_GAIN: keep = in_db >= thr && in_db > -100.0f; v = r_out[i] - in_db; break;
                case CM_SILENCE_DELTA: keep = in_db < thr && in_db > -100.0f; v = r_out[i] - in_db; break;
                default: keep = in_db < thr; v = -fmaxf(r_val[i], 0.0f); break;
            }
            if (keep) dst[c++] = v;
        }
        sh[31] = (float)c;
    }
    co.sync();
    return (int)sh[31];
}

AF_HD void body_finalize(const BatchArgs& a, int s, const Coop& co, float* ws) {
    const size_t stride = (size_t)a.stride;
    const int n_rows = a.n_rows;
    float* r_in = ws;
    float* r_out = r_in + n_rows;
    float* r_comp = r_out + n_rows;
    float* r_de = r_comp + n_rows;
    float* buf = r_de + n_rows;
    float* buf2 = buf + a.n_pad;
    float* sh = buf2 + a.n_pad;
    for (int i = co.tid; i < n_rows; i += co.n) {
        r_in[i] = a.rows[((size_t)0 * n_rows + i) * stride + s];
        r_out[i] = a.rows[((size_t)1 * n_rows + i) * stride + s];
        r_comp[i] = a.rows[((size_t)2 * n_rows + i) * stride + s];
        r_de[i] = a.rows[((size_t)3 * n_rows + i) * stride + s];
        buf[i] = r_in[i];
    }
    co.sync();
    // input RMS rows -> p20 / p90 -> active threshold (:591-596)
    coop_sort(co, buf, n_rows);
    if (co.tid == 0) {
        const float floor_db = sorted_percentile(buf, n_rows, 0.20f);
        const float p90 = sorted_percentile(buf, n_rows, 0.90f);
        sh[0] = fmaxf(fmaxf(floor_db + 6.0f, p90 - 24.0f), -60.0f);
    }
    co.sync();
    const float thr = sh[0];

    // active compressor / de-esser gain reduction (:597-625)
    int n_act = coop_compact(co, buf, n_rows, r_in, r_comp, nullptr, CM_ACTIVE_GR, thr, sh);
    int mode = CM_ACTIVE_GR;
    if (n_act < 3) {
        mode = CM_ALL_GR;
        n_act = coop_compact(co, buf, n_rows, r_in, r_comp, nullptr, CM_ALL_GR, thr, sh);
    }
    if (co.tid == 0) {
        int c = 0;
        for (int i = 0; i < n_act; ++i) c += buf[i] >= 0.10f ? 1 : 0;
        sh[1] = n_act > 0 ? (float)c / (float)n_act : 0.0f;
    }
    co.sync();
    coop_sort(co, buf, n_act);
    if (co.tid == 0) {
        sh[2] = sorted_percentile(buf, n_act, 0.50f);
        sh[3] = sorted_percentile(buf, n_act, 0.95f);
    }
    co.sync();
    const int n_de = coop_compact(co, buf, n_rows, r_in, r_de, nullptr, mode, thr, sh);
    coop_sort(co, buf, n_de);
    if (co.tid == 0) {
        sh[4] = sorted_percentile(buf, n_de, 0.50f);
        sh[5] = sorted_percentile(buf, n_de, 0.95f);
    }
    co.sync();
    // active output gain, silence level delta, silence output gain: medians (:626-648)
    for (int q = 0; q < 3; ++q) {
        const int m = q == 0 ? CM_ACTIVE_GAIN : (q == 1 ? CM_SILENCE_DELTA : CM_SILENCE_GAIN);
        const int cnt = coop_compact(co, buf, n_rows, r_in, r_comp, r_out, m, thr, sh);
        coop_sort(co, buf, cnt);
        if (co.tid == 0) sh[6 + q] = sorted_percentile(buf, cnt, 0.50f);
        co.sync();
    }
    // compressor pumping score (:74-111) over all rows' clamped GR at 50 Hz
    float pumping = 0.0f;
    if (n_rows >= 3) {
        if (co.tid == 0) {
            const float dt = 1.0f / 50.0f;
            const float pi = 3.14159265358979323846f;
            const float hp_rc = 1.0f / (2.0f * pi * 2.0f);
            const float lp_rc = 1.0f / (2.0f * pi * 8.0f);
            const float hp_alpha = hp_rc / (hp_rc + dt);
            const float lp_alpha = dt / (lp_rc + dt);
            float prev = fmaxf(r_comp[0], 0.0f), hp = 0.0f, bp = 0.0f;
            int bad = 0;
            for (int i = 1; i < n_rows; ++i) {
                const float v = fmaxf(r_comp[i], 0.0f);
                if (!af_finite(v)) {
                    bad = 1;
                    break;
                }
                hp = hp_alpha * (hp + v - prev);
                bp += lp_alpha * (hp - bp);
                buf[i - 1] = fabsf(bp);
                buf2[i - 1] = fabsf(v - prev);
                prev = v;
            }
            sh[10] = (float)bad;
        }
        co.sync();
        if (sh[10] != 0.0f) {
            pumping = INFINITY;
        } else {
            const int m = n_rows - 1;
            coop_sort(co, buf2, m);  // p95 of |delta|
            if (co.tid == 0) sh[11] = sorted_percentile(buf2, m, 0.95f);
            co.sync();
            for (int i = co.tid; i < m; i += co.n) buf2[i] = buf[i];
            co.sync();
            coop_sort(co, buf2, m);  // clip limit = p95 of |band-passed trace|; buf keeps trace order
            if (co.tid == 0) {
                const float limit = sorted_percentile(buf2, m, 0.95f);
                float sum = 0.0f;
                for (int i = 0; i < m; ++i) {
                    const float v = fminf(buf[i], limit);
                    sum += v * v;
                }
                sh[12] = sqrtf(sum / (float)m) + sh[11];
            }
            co.sync();
            pumping = sh[12];
        }
    }
    if (co.tid == 0) {
        float max_comp = 0.0f, max_de = 0.0f;
        for (int i = 0; i < n_rows; ++i) {
            max_comp = fmaxf(max_comp, r_comp[i]);
            max_de = fmaxf(max_de, r_de[i]);
        }
        const StreamAccum acc = a.accum[s];
        const float ceiling = stream_params(a, s).effective_ceiling_db;
        const int T = a.n_samples;
        const float in_rms = T ? (float)sqrt(acc.sum_in / (double)T) : 0.0f;
        const float out_rms = T ? (float)sqrt(acc.sum_out / (double)T) : 0.0f;
        AfChainMetrics m;
        m.input_sample_peak_db = lin_to_db_f32(acc.peak_in);
        m.input_rms_db = lin_to_db_f32(in_rms);
        m.output_sample_peak_db = lin_to_db_f32(acc.peak_out);
        m.pre_limiter_true_peak_db = lin_to_db_f32(acc.peak_pre_tp);
        m.output_true_peak_db = lin_to_db_f32(acc.peak_out_tp);
        m.output_rms_db = lin_to_db_f32(out_rms);
        m.limiter_effective_ceiling_db = ceiling;
        m.sample_headroom_db = ceiling - m.output_sample_peak_db;
        m.pre_limiter_true_peak_headroom_db = ceiling - m.pre_limiter_true_peak_db;
        m.true_peak_headroom_db = ceiling - m.output_true_peak_db;
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
        m.processed_samples = (uint64_t)T;
        m.candidate_runtime_ms = 0.0;
        static_cast<AfChainMetrics*>(a.metrics)[a.pair[s]] = m;
    }
    co.sync();
}

}  // namespace afsim
