// afsim_eqscan.h -- time-parallel biquad cascade for ONE long passage (simulate_eq_v2 on a single capture).
//
// A batch of one stream gives the stage kernels nothing to parallelise over, and a DF2T section is a serial
// recurrence.  It is also linear: with the state z = (z1, z2),
//     y = b0 x + z1,   z' = A z + B x,   A = [[-a1, 1], [-a2, 0]],   B = (b1 - a1 b0, b2 - a2 b0)
// (dsp/biquad.rs:262-274), so the passage is cut into P segments of L samples and each section runs in three steps:
//   local   every segment is filtered from a ZERO state; its end state e_k is the forced response
//   scan    the true state at the start of segment k follows s_(k+1) = A^L s_k + e_k: an inclusive scan of the
//           affine maps (A^L, e_k) -- 2x2 matrices and 2-vectors combined with warp shuffles (Kogge-Stone inside a
//           warp, then across the warps of one block through shared memory)
//   apply   every segment is filtered again from its true start state s_k: the ordinary recurrence, so every
//           sample is produced by the reference's own DF2T arithmetic; only s_k carries the scan's rounding
// `apply` of section j and `local` of section j+1 share one pass over the data.  The samples between sections are
// rounded to f32 exactly as in the serial cascade (dsp/biquad.rs:262 `as f32`).  Results differ from the serial
// walk by the rounding of s_k (~1e-15 relative), far inside the 1e-5 / -100 dBFS render tolerance; the batched
// path (afsim_stages.h EqStage) stays the bit-exact one.
#pragma once
#include "afsim_stages.h"

namespace afsim {

struct Affine2 {  // z -> M z + v
    double m00, m01, m10, m11, v0, v1;
};
AF_HD Affine2 affine_identity() { return Affine2{1.0, 0.0, 0.0, 1.0, 0.0, 0.0}; }
// later o earlier: apply `a` first, then `b`
AF_HD Affine2 affine_then(const Affine2& a, const Affine2& b) {
    Affine2 r;
    r.m00 = b.m00 * a.m00 + b.m01 * a.m10;
    r.m01 = b.m00 * a.m01 + b.m01 * a.m11;
    r.m10 = b.m10 * a.m00 + b.m11 * a.m10;
    r.m11 = b.m10 * a.m01 + b.m11 * a.m11;
    r.v0 = b.m00 * a.v0 + b.m01 * a.v1 + b.v0;
    r.v1 = b.m10 * a.v0 + b.m11 * a.v1 + b.v1;
    return r;
}
// A^(2^log2_len) of a section by repeated squaring (segment lengths are powers of two)
AF_HD void biquad_transition_power(const Bq& c, int log2_len, double (&m)[4]) {
    double a00 = -c.a1, a01 = 1.0, a10 = -c.a2, a11 = 0.0;
    for (int i = 0; i < log2_len; ++i) {
        const double n00 = a00 * a00 + a01 * a10;
        const double n01 = a00 * a01 + a01 * a11;
        const double n10 = a10 * a00 + a11 * a10;
        const double n11 = a10 * a01 + a11 * a11;
        a00 = n00;
        a01 = n01;
        a10 = n10;
        a11 = n11;
    }
    m[0] = a00;
    m[1] = a01;
    m[2] = a10;
    m[3] = a11;
}

// One pass over segment `seg` of the time-major-transposed signal xt[i * n_seg + seg], i = 0 .. len-1:
//   APPLY: section `cur` runs from its true start state (s1, s2) and overwrites the samples with its f32 output;
//   LOCAL: section `nxt` runs on that output from a zero state; its end state goes to (e1, e2).
template <bool APPLY, bool LOCAL>
AF_HD void eqscan_segment(float* xt, size_t n_seg, size_t seg, int len, const Bq& cur, double s1, double s2, const Bq& nxt,
                          double* e1, double* e2) {
    double z1 = s1, z2 = s2, n1 = 0.0, n2 = 0.0;
    float* p = xt + seg;
    constexpr int U = 8;  // tiles of 8: the loads of a tile are issued together (segment lengths are multiples of 8)
    int i = 0;
    for (; i + U <= len; i += U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = p[(size_t)(i + u) * n_seg];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (APPLY) v[u] = (float)bq_step((double)v[u], cur, z1, z2);
            if (LOCAL) (void)bq_step((double)v[u], nxt, n1, n2);
        }
        if (APPLY) {
#pragma unroll
            for (int u = 0; u < U; ++u) p[(size_t)(i + u) * n_seg] = v[u];
        }
    }
    for (; i < len; ++i) {
        float v = p[(size_t)i * n_seg];
        if (APPLY) {
            v = (float)bq_step((double)v, cur, z1, z2);
            p[(size_t)i * n_seg] = v;
        }
        if (LOCAL) (void)bq_step((double)v, nxt, n1, n2);
    }
    if (LOCAL) {
        *e1 = n1;
        *e2 = n2;
    }
}

}  // namespace afsim
