// afsim_cleanup.h -- adaptive input cleanup (hum / harmonic notches, rumble-adaptive high-pass) as an
// optional input stage of the chain render.
//
// Restates AdaptiveInputCleanupState of the reference's live loop
// (rust-core/src/audio/processor/routing.rs:55-648) with the block contract of its test harness
// (processor/tests.rs:500-549): per 480-sample block `analyze_input(raw block)` -> DC block ->
// `process_block(block)`.  Everything is f32 in the reference's evaluation order (no contraction);
// only the window-end transcendentals (atan2f / logf) and the notch retune (sinf / cosf) go through
// the device libm, so renders agree with the oracle to float rounding of those few values.
//
// 13 + 13 rotating-phasor bins (49..61 Hz and their second harmonics) are 26 independent chains per
// sample: one thread per stream keeps them in registers and the scheduler overlaps them.
#pragma once
#include "afsim_stages.h"

namespace afsim {

constexpr int kHumBins = 13;

// Sample-rate constants, derived on the host with the host libm exactly as the constructors do.
struct CleanupConst {
    float fs;
    float lowpass_coeff;          // clamp(2*pi*150/fs, 0, 1)          routing.rs:341
    float bin_cos[2 * kHumBins];  // HumBin::new: cos / sin of 2*pi*f/fs; [0..12] primary, [13..25] harmonic
    float bin_sin[2 * kHumBins];
    float notch0[2][5];           // NotchFilter::new(55 Hz / 110 Hz, Q 36)   routing.rs:117-140
    double hp[3][5];              // adaptive high-pass at 80 / 100 / 120 Hz, Q 0.707
    int window_samples;           // round(fs * 0.25)
    int notch_fade_total;         // round(fs * 0.020)
    int hp_fade_total;            // biquad crossfade F
    uint32_t rumble_hold_gentle;  // round(fs * 0.18)
    uint32_t rumble_hold_strong;  // round(fs * 0.30)
    uint32_t hum_hold;            // round(fs * 0.75)
};

struct NotchF32 {  // routing.rs:117-157
    float b0, b1, b2, a1, a2, z1, z2;
    AF_HD float process(float x) {
        const float y = b0 * x + z1;
        z1 = b1 * x - a1 * y + z2;
        z2 = b2 * x - a2 * y;
        return y;
    }
    AF_HD void design(float f, float q, float fs) {
        const float pi = 3.14159265358979323846f;
        const float omega = 2.0f * pi * f / fmaxf(fs, 1.0f);
        const float sn = sinf(omega), cs = cosf(omega);
        const float alpha = sn / (2.0f * fmaxf(q, 1.0f));
        const float a0 = 1.0f + alpha;
        b0 = 1.0f / a0;
        b1 = -2.0f * cs / a0;
        b2 = 1.0f / a0;
        a1 = -2.0f * cs / a0;
        a2 = (1.0f - alpha) / a0;
        z1 = z2 = 0.0f;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
        io.f32(b0);
        io.f32(b1);
        io.f32(b2);
        io.f32(a1);
        io.f32(a2);
        io.f32(z1);
        io.f32(z2);
    }
};

struct SmoothNotchF32 {  // routing.rs:160-217
    NotchF32 active, pending;
    float frequency_hz, pending_frequency_hz;
    int fade_remaining;
    AF_HD void init(const float* coeffs, float f) {
        active.b0 = coeffs[0];
        active.b1 = coeffs[1];
        active.b2 = coeffs[2];
        active.a1 = coeffs[3];
        active.a2 = coeffs[4];
        active.z1 = active.z2 = 0.0f;
        pending = active;
        frequency_hz = pending_frequency_hz = f;
        fade_remaining = 0;
    }
    AF_HD void retune(float f, float fs, int fade_total) {
        f = clampf(f, 20.0f, fs * 0.45f);
        if (fabsf(f - pending_frequency_hz) < 0.15f) return;
        pending.design(f, 36.0f, fs);
        pending_frequency_hz = f;
        fade_remaining = fade_total;
    }
    AF_HD float process(float x, int fade_total) {
        const float a = active.process(x);
        if (fade_remaining == 0) return a;
        const float p = pending.process(x);
        const float fade = (float)(fade_total - fade_remaining + 1) / (float)fade_total;
        const float out = a + (p - a) * fade;
        fade_remaining -= 1;
        if (fade_remaining == 0) {
            active = pending;
            frequency_hz = pending_frequency_hz;
        }
        return out;
    }
    template <class IO>
    AF_HD void sync(IO& io) {
        active.sync(io);
        pending.sync(io);
        io.f32(frequency_hz);
        io.f32(pending_frequency_hz);
        int32_t fr = fade_remaining;
        uint32_t u = (uint32_t)fr;
        io.u32(u);
        fade_remaining = (int)u;
    }
};

AF_HD float wrap_phase_f32(float p) {  // routing.rs:598-607
    const float pi = 3.14159265358979323846f;
    const float two_pi = 2.0f * pi;
    while (p > pi) p -= two_pi;
    while (p < -pi) p += two_pi;
    return p;
}
AF_HD float smooth_toward_f32(float cur, float target, float attack, float release) {  // routing.rs:642-646
    const float c = target > cur ? attack : release;
    return cur + c * (target - cur);
}

// NB = 26: one thread owns all bins of its stream (the per-stream kernels and the host harness).
// NB = 1: one WARP per stream, lane i owns bin i (the shared input stage of a candidate sweep renders a handful of
// distinct passages, so its serial chain is the wavefront's critical path: spreading the 26 independent phasor
// chains over the lanes cuts the instructions per sample from ~490 to ~130).  Everything that is not a bin is
// computed redundantly -- and identically -- by every lane; the window decision gathers the bins with shuffles.
template <int NB>
struct CleanupStageT {
    // hum analyser: rotating phasors + accumulators of this thread's bins [bin0, bin0 + NB)
    float cosp[NB], sinp[NB], iacc[NB], qacc[NB];
    int bin0;
    float my_bin_cos, my_bin_sin;  // NB == 1: this lane's rotation (CleanupConst::bin_cos / bin_sin of bin0)
    float lowpass, low_env, slow_low_env, broadband_env, total_energy;
    float hum_line_hz, prev_phase, hum_strength, harmonic_strength, highpass_hz;
    uint32_t rumble_hold, hum_hold, windows_observed, window_pos, candidate_windows;
    bool phase_valid;
    // adaptive high-pass: f64 biquad with the 1.5 ms coefficient crossfade (dsp/biquad.rs:249-327)
    Bq hp_active, hp_pending;
    double hz1, hz2, hpz1, hpz2;
    uint32_t hp_fade_remaining;
    SmoothNotchF32 hum_notch, harmonic_notch;

    AF_HD void init(const CleanupConst& k, int first_bin = 0) {
        bin0 = first_bin;
        my_bin_cos = k.bin_cos[first_bin];
        my_bin_sin = k.bin_sin[first_bin];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            cosp[i] = 1.0f;
            sinp[i] = 0.0f;
            iacc[i] = qacc[i] = 0.0f;
        }
        lowpass = low_env = slow_low_env = broadband_env = total_energy = 0.0f;
        hum_line_hz = prev_phase = hum_strength = harmonic_strength = 0.0f;
        highpass_hz = 80.0f;
        rumble_hold = hum_hold = windows_observed = window_pos = candidate_windows = 0;
        phase_valid = false;
        hp_active = bq_from(k.hp[0]);
        hp_pending = hp_active;
        hz1 = hz2 = hpz1 = hpz2 = 0.0;
        hp_fade_remaining = 0;
        hum_notch.init(k.notch0[0], 55.0f);
        harmonic_notch.init(k.notch0[1], 110.0f);
    }

    template <class IO>
    AF_HD void sync(IO& io) {
        if (NB == 2 * kHumBins) {
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                io.f32(cosp[i]);
                io.f32(sinp[i]);
                io.f32(iacc[i]);
                io.f32(qacc[i]);
            }
        } else {  // same table layout: bin b lives in slots 4b .. 4b+3
            IO mine = io;
            mine.p += (size_t)(4 * bin0) * io.stride;
            mine.f32(cosp[0]);
            mine.f32(sinp[0]);
            mine.f32(iacc[0]);
            mine.f32(qacc[0]);
            io.p += (size_t)(4 * 2 * kHumBins) * io.stride;
        }
        io.f32(lowpass);
        io.f32(low_env);
        io.f32(slow_low_env);
        io.f32(broadband_env);
        io.f32(total_energy);
        io.f32(hum_line_hz);
        io.f32(prev_phase);
        io.f32(hum_strength);
        io.f32(harmonic_strength);
        io.f32(highpass_hz);
        io.u32(rumble_hold);
        io.u32(hum_hold);
        io.u32(windows_observed);
        io.u32(window_pos);
        io.u32(candidate_windows);
        io.flag(phase_valid);
        io.f64(hp_active.b0);
        io.f64(hp_active.b1);
        io.f64(hp_active.b2);
        io.f64(hp_active.a1);
        io.f64(hp_active.a2);
        io.f64(hp_pending.b0);
        io.f64(hp_pending.b1);
        io.f64(hp_pending.b2);
        io.f64(hp_pending.a1);
        io.f64(hp_pending.a2);
        io.f64(hz1);
        io.f64(hz2);
        io.f64(hpz1);
        io.f64(hpz2);
        io.u32(hp_fade_remaining);
        hum_notch.sync(io);
        harmonic_notch.sync(io);
    }

    // routing.rs:405-532: one finished 250 ms window
    AF_HD void finish_window(const CleanupConst& k, bool gentle) {
        const float pi = 3.14159265358979323846f;
        const float n = (float)(k.window_samples > 1 ? k.window_samples : 1);
        float primary[kHumBins], harmonic[kHumBins], phase[kHumBins];
        float best_f = 0.0f, best_primary = 0.0f, best_harm = 0.0f, best_score = 0.0f, best_phase = 0.0f;
        float my_power[NB], my_phase[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {  // HumBin::power_phase_and_reset, routing.rs:90-109
            my_power[i] = (iacc[i] * iacc[i] + qacc[i] * qacc[i]) * (2.0f / (n * n));
            my_phase[i] = atan2f(qacc[i], iacc[i]);
            iacc[i] = 0.0f;
            qacc[i] = 0.0f;
            const float norm = sqrtf(cosp[i] * cosp[i] + sinp[i] * sinp[i]);
            if (norm > 1.0e-6f) {
                cosp[i] /= norm;
                sinp[i] /= norm;
            }
        }
        if (NB == 2 * kHumBins) {
#pragma unroll
            for (int i = 0; i < kHumBins; ++i) {
                primary[i] = my_power[i % NB];
                phase[i] = my_phase[i % NB];
                harmonic[i] = my_power[(kHumBins + i) % NB];
            }
        } else {
#if defined(__CUDA_ARCH__)
#pragma unroll
            for (int i = 0; i < kHumBins; ++i) {  // every lane ends up with all 26 bins
                primary[i] = __shfl_sync(0xffffffffu, my_power[0], i);
                phase[i] = __shfl_sync(0xffffffffu, my_phase[0], i);
                harmonic[i] = __shfl_sync(0xffffffffu, my_power[0], kHumBins + i);
            }
#else
#pragma unroll
            for (int i = 0; i < kHumBins; ++i) primary[i] = phase[i] = harmonic[i] = 0.0f;  // the lane form is device only
#endif
        }
#pragma unroll
        for (int i = 0; i < kHumBins; ++i) {
            const float score = primary[i] + harmonic[i] * 0.65f;
            if (score > best_score) {
                best_score = score;
                best_primary = primary[i];
                best_harm = harmonic[i];
                best_f = 49.0f + (float)i * 1.0f;
                best_phase = phase[i];
            }
        }
        const float total_power = total_energy / n + 1.0e-9f;
        window_pos = 0;
        windows_observed = windows_observed == 0xffffffffu ? windows_observed : windows_observed + 1;
        total_energy = 0.0f;
        const float primary_ratio = best_primary / total_power;
        const float harmonic_ratio = best_harm / total_power;
        const float ratio_threshold = gentle ? 0.075f : 0.040f;
        const float power_threshold = gentle ? 1.8e-5f : 8.0e-6f;
        const bool candidate = (best_primary > power_threshold || best_harm > power_threshold * 0.70f) &&
                               (primary_ratio > ratio_threshold || harmonic_ratio > ratio_threshold * 0.85f) && best_f > 0.0f;
        if (candidate) {
            candidate_windows = candidate_windows + 1 < 3 ? candidate_windows + 1 : 3;
        } else {
            candidate_windows = 0;
            phase_valid = false;
        }
        if (candidate_windows >= 2) {
            hum_hold = k.hum_hold;
            const float idx_f = clampf(roundf((best_f - 49.0f) / 1.0f), 0.0f, 12.0f);
            const int idx = (int)idx_f;
            float offset = 0.0f;
            if (idx > 0 && idx + 1 < kHumBins) {
                float left = 0.0f, center = 0.0f, right = 0.0f;
#pragma unroll
                for (int i = 0; i < kHumBins; ++i) {  // dynamic index without a local-memory array
                    if (i == idx - 1) left = primary[i];
                    if (i == idx) center = primary[i];
                    if (i == idx + 1) right = primary[i];
                }
                left = logf(fmaxf(left, 1.0e-12f));
                center = logf(fmaxf(center, 1.0e-12f));
                right = logf(fmaxf(right, 1.0e-12f));
                const float denom = left - 2.0f * center + right;
                if (fabsf(denom) > 1.0e-6f) offset = clampf(0.5f * (left - right) / denom, -0.5f, 0.5f);
            }
            const float spectral = clampf(best_f + offset * 1.0f, 49.0f, 61.0f);
            const float fs1 = fmaxf(k.fs, 1.0f);
            const float window_seconds = (float)k.window_samples / fs1;
            const float center_sample = ((float)windows_observed + 0.5f) * (float)k.window_samples;
            const float absolute_phase = wrap_phase_f32(-best_phase + 2.0f * pi * best_f * center_sample / fs1);
            bool have_phase = false;
            float phase_hz = 0.0f;
            if (phase_valid && window_seconds > 0.0f) {
                const float delta = wrap_phase_f32(absolute_phase - prev_phase);
                const float base = delta / (2.0f * pi * window_seconds);
                const float spacing = 1.0f / window_seconds;
                float best_alias = base, best_err = INFINITY;
                for (int a = -32; a <= 32; ++a) {
                    const float cand = base + (float)a * spacing;
                    const float err = fabsf(cand - spectral);
                    if (err < best_err) {
                        best_alias = cand;
                        best_err = err;
                    }
                }
                phase_hz = clampf(best_alias, 49.0f, 61.0f);
                have_phase = true;
            }
            const float measured = have_phase ? 0.75f * spectral + 0.25f * phase_hz : spectral;
            const float next = hum_line_hz <= 0.0f ? measured : hum_line_hz + 0.35f * (measured - hum_line_hz);
            hum_line_hz = clampf(next, 49.0f, 61.0f);
            prev_phase = absolute_phase;
            phase_valid = true;
        }
    }

    // routing.rs:336-403: one raw input sample
    AF_HD void analyze(float s, const CleanupConst& k, bool gentle) {
        total_energy += s * s;
#pragma unroll
        for (int i = 0; i < NB; ++i) {  // HumBin::analyze, routing.rs:78-88
            const float bc = NB == 2 * kHumBins ? k.bin_cos[i] : my_bin_cos, bs = NB == 2 * kHumBins ? k.bin_sin[i] : my_bin_sin;
            iacc[i] += s * cosp[i];
            qacc[i] += s * sinp[i];
            const float nc = cosp[i] * bc - sinp[i] * bs;
            const float ns = sinp[i] * bc + cosp[i] * bs;
            cosp[i] = nc;
            sinp[i] = ns;
        }
        window_pos += 1;
        if (window_pos >= (uint32_t)k.window_samples) finish_window(k, gentle);
        lowpass += k.lowpass_coeff * (s - lowpass);
        const float low_abs = fabsf(lowpass);
        const float lc = low_abs > low_env ? 0.08f : 0.006f;
        low_env += lc * (low_abs - low_env);
        slow_low_env += 0.0012f * (low_abs - slow_low_env);
        broadband_env += 0.02f * (fabsf(s) - broadband_env);
        const float burst_ratio = low_env / fmaxf(slow_low_env, 0.006f);
        const float low_dominance = low_env / fmaxf(broadband_env, 0.01f);
        const float threshold = gentle ? 0.055f : 0.035f;
        const float ratio_threshold = gentle ? 2.8f : 2.1f;
        const bool startup_burst = windows_observed == 0 && low_env > 0.45f;
        const bool established = windows_observed > 0 && slow_low_env > 0.012f;
        if ((startup_burst || established) && hum_hold == 0 && candidate_windows == 0 && low_env > threshold &&
            burst_ratio > ratio_threshold && low_dominance > 0.62f) {
            rumble_hold = gentle ? k.rumble_hold_gentle : k.rumble_hold_strong;
        } else {
            rumble_hold = rumble_hold > 0 ? rumble_hold - 1 : 0;
        }
        hum_hold = hum_hold > 0 ? hum_hold - 1 : 0;
    }

    // routing.rs:534-585: decisions taken once per block, before its samples are filtered
    AF_HD void begin_block(const CleanupConst& k, bool gentle, bool* hum_detected) {
        *hum_detected = hum_hold > 0;
        const bool rumble_detected = rumble_hold > 0;
        const float selected = rumble_detected ? (gentle ? 100.0f : 120.0f) : 80.0f;
        if (fabsf(selected - highpass_hz) > 0.5f) {  // Biquad::set_frequency -> coefficient crossfade
            hp_pending = bq_from(k.hp[selected == 80.0f ? 0 : (selected == 100.0f ? 1 : 2)]);
            hpz1 = hz1;
            hpz2 = hz2;
            hp_fade_remaining = (uint32_t)k.hp_fade_total;
            highpass_hz = selected;
        }
        const float attack = gentle ? 0.22f : 0.34f;
        const float release = 0.035f;
        const float target_hum = *hum_detected ? (gentle ? 0.55f : 0.85f) : 0.0f;
        const float target_harm = *hum_detected ? (gentle ? 0.0f : 0.60f) : 0.0f;
        hum_strength = smooth_toward_f32(hum_strength, target_hum, attack, release);
        harmonic_strength = smooth_toward_f32(harmonic_strength, target_harm, attack, release);
        if (hum_line_hz > 0.0f) {
            hum_notch.retune(hum_line_hz, k.fs, k.notch_fade_total);
            harmonic_notch.retune(hum_line_hz * 2.0f, k.fs, k.notch_fade_total);
        }
    }

    // routing.rs:587-595 + dsp/biquad.rs:290-327
    AF_HD float process(float y, const CleanupConst& k) {
        const float pn = hum_notch.process(y, k.notch_fade_total);
        y += (pn - y) * clampf(hum_strength, 0.0f, 1.0f);
        const float hn = harmonic_notch.process(y, k.notch_fade_total);
        y += (hn - y) * clampf(harmonic_strength, 0.0f, 1.0f);
        const double x = (double)y;
        const double ya = bq_step(x, hp_active, hz1, hz2);
        if (hp_fade_remaining == 0) return (float)ya;
        const double yp = bq_step(x, hp_pending, hpz1, hpz2);
        const uint32_t total = (uint32_t)k.hp_fade_total;
        const double fade = (double)(total - hp_fade_remaining + 1) / (double)total;
        const double out = ya * (1.0 - fade) + yp * fade;
        hp_fade_remaining -= 1;
        if (hp_fade_remaining == 0) {
            hp_active = hp_pending;
            hz1 = hpz1;
            hz2 = hpz2;
        }
        return (float)out;
    }
};

using CleanupStage = CleanupStageT<2 * kHumBins>;

}  // namespace afsim
