// afsim_plan.cpp -- host planner (see afsim_plan.h).
//
// Walks the reference's constructor + setter sequence for one candidate and leaves only the
// constants the device recurrences need.  All transcendental work that does not depend on the
// signal (time-constant exponentials, dB conversions of fixed settings, RBJ coefficient design)
// happens here with the host libm, exactly where the reference evaluates it, so those values are
// bit-identical to a CPU render.  Build with -ffp-contract=off.
//
// Citations are relative to rust-core/src/.
#include "afsim_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <set>
#include <utility>

namespace afsim {

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;

// Rust f64::max / min ignore a NaN operand; clamp keeps NaN.
inline double rmax(double a, double b) { return std::fmax(a, b); }
inline double rmin(double a, double b) { return std::fmin(a, b); }
inline double rclamp(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline float rclampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline size_t as_usize(double v) {  // `as usize`: saturating, NaN -> 0
    if (!(v > 0.0)) return 0;
    if (v >= 18446744073709551615.0) return std::numeric_limits<size_t>::max();
    return static_cast<size_t>(v);
}
inline double db_to_linear(double db) { return std::pow(10.0, db / 20.0); }  // dsp/util.rs:11-14
inline double lerp(double a, double b, double t) { return a + (b - a) * t; }

int fail(std::string* error, int status, const std::string& message) {
    if (error) *error = message;
    return status;
}

std::string fmt_num(double v) {
    // Rust `{}` for f64: the shortest digits that round-trip, always in positional notation.
    char buf[400];
    if (!std::isfinite(v)) {
        std::snprintf(buf, sizeof buf, "%s", std::isnan(v) ? "NaN" : (v > 0 ? "inf" : "-inf"));
        return buf;
    }
    for (int decimals = 0; decimals <= 340; ++decimals) {
        std::snprintf(buf, sizeof buf, "%.*f", decimals, v);
        if (std::strtod(buf, nullptr) == v) break;
    }
    return buf;
}

void store(double dst[5], const BiquadCoeffs& c) {
    dst[0] = c.b0;
    dst[1] = c.b1;
    dst[2] = c.b2;
    dst[3] = c.a1;
    dst[4] = c.a2;
}
void store_identity(double dst[5]) {
    dst[0] = 1.0;
    dst[1] = dst[2] = dst[3] = dst[4] = 0.0;
}

// dsp/eq.rs:46-53 EqFilterType ids -> biquad kind
BqKind kind_of(uint8_t filter_type) {
    switch (filter_type) {
        case AF_LOW_SHELF: return BqKind::LowShelf;
        case AF_BELL: return BqKind::Peaking;
        case AF_HIGH_SHELF: return BqKind::HighShelf;
        case AF_NOTCH: return BqKind::Notch;
        case AF_HIGH_PASS: return BqKind::HighPass;
        default: return BqKind::LowPass;
    }
}
bool is_pass(uint8_t t) { return t == AF_HIGH_PASS || t == AF_LOW_PASS; }
bool supported_slope(uint8_t s) { return s == 12 || s == 24 || s == 36 || s == 48; }

double butterworth_section_q(size_t index, size_t count) {  // dsp/eq.rs:203-207
    const size_t order = 2 * count;
    const double angle = static_cast<double>(2 * index + 1) * kPi / static_cast<double>(2 * order);
    return 1.0 / (2.0 * std::cos(angle));
}

// dsp/eq.rs:154-201; "" when valid
std::string validate_band(const AfBand& b, size_t index, double fs) {
    const std::string prefix = "Band " + std::to_string(index) + ": ";
    if (!std::isfinite(b.frequency_hz)) return prefix + "frequency must be finite";
    if (!std::isfinite(fs) || fs <= 40.0) return prefix + "sample rate must be finite and support the EQ frequency range";
    const double max_f = rmax(fs / 2.0 - 1.0, 20.0);
    if (!(b.frequency_hz >= 20.0 && b.frequency_hz <= max_f))
        return prefix + "frequency " + fmt_num(b.frequency_hz) + " Hz out of range [20, " + fmt_num(max_f) + "]";
    if (!std::isfinite(b.gain_db)) return prefix + "gain must be finite";
    if (!(b.gain_db >= -12.0 && b.gain_db <= 12.0))
        return prefix + "gain " + fmt_num(b.gain_db) + " dB out of range [-12, 12]";
    if (!std::isfinite(b.q)) return prefix + "Q must be finite";
    if (!(b.q >= 0.1 && b.q <= 10.0)) return prefix + "Q " + fmt_num(b.q) + " out of range [0.1, 10]";
    if (!supported_slope(b.slope_db_per_octave))
        return prefix + "slope " + std::to_string(b.slope_db_per_octave) +
               " dB/octave is unsupported; expected one of [12, 24, 36, 48]";
    return "";
}

constexpr double kDefaultFrequencies[10] = {80.0, 160.0, 320.0, 640.0, 1280.0, 2500.0, 5000.0, 8000.0, 12000.0, 16000.0};
constexpr double kDefaultQ = 1.41;
BqKind default_kind(size_t band) {  // dsp/eq.rs:125-140
    return band == 0 ? BqKind::LowShelf : (band == 9 ? BqKind::HighShelf : BqKind::Peaking);
}

// Sections of the configured EQ, flattened band-major (dsp/eq.rs:258-277, 340-350).
// Returns the section count; band_sections (nullable) gets the per-band counts.
int design_eq(const AfBand bands[AFSIM_NUM_BANDS], bool typed, double fs, double out[kMaxSections][5], bool band_major_slots,
              int* band_sections) {
    int n = 0;
    for (size_t i = 0; i < AFSIM_NUM_BANDS; ++i) {
        const AfBand& b = bands[i];
        int count = 0;
        if (!typed) {
            // legacy setters keep the default band's type and enabled flag (python_api.rs:408-412)
            const int slot = band_major_slots ? static_cast<int>(i) * 4 : n;
            store(out[slot], design_biquad(default_kind(i), b.frequency_hz, b.gain_db, b.q, fs));
            count = 1;
        } else if (b.enabled) {
            if (is_pass(b.filter_type)) {
                count = supported_slope(b.slope_db_per_octave) ? b.slope_db_per_octave / 12 : 1;
                for (int k = 0; k < count; ++k) {
                    const int slot = band_major_slots ? static_cast<int>(i) * 4 + k : n + k;
                    store(out[slot], design_biquad(kind_of(b.filter_type), b.frequency_hz, 0.0,
                                                   butterworth_section_q(static_cast<size_t>(k), static_cast<size_t>(count)), fs));
                }
            } else {
                const double gain = b.filter_type == AF_NOTCH ? 0.0 : b.gain_db;  // dsp/eq.rs:270-274
                const BiquadCoeffs c = design_biquad(kind_of(b.filter_type), b.frequency_hz, gain, b.q, fs);
                // A flat band (0 dB bell / shelf) designs to b0 == 1, b1 == a1, b2 == a2: DF2T then returns
                // its input bit for bit and keeps z1 = z2 = 0 for every finite sample, so the render cascade
                // drops the section (the typed path has no coefficient crossfade that could tell the difference).
                const bool identity = c.b0 == 1.0 && c.b1 == c.a1 && c.b2 == c.a2;
                if (band_major_slots || !identity) {
                    const int slot = band_major_slots ? static_cast<int>(i) * 4 : n;
                    store(out[slot], c);
                    count = 1;
                }
            }
        }
        if (band_sections) band_sections[i] = count;
        n += count;
    }
    return n;
}

// processor/control.rs:904-910
double effective_limiter_ceiling_db(double ceiling_db, bool careful) { return careful ? rmin(ceiling_db, -1.5) : ceiling_db; }

bool loudness_meter_supports(double fs) {  // dsp/loudness.rs:37-42 with `sample_rate as u32`
    const uint32_t rate = static_cast<uint32_t>(std::min<size_t>(as_usize(fs), 0xffffffffu));
    const uint32_t ok[] = {8000, 16000, 32000, 44100, 48000, 88200, 96000};
    return std::find(std::begin(ok), std::end(ok), rate) != std::end(ok);
}

bool plan_deesser(const AfChainSettings& s, double fs, CandidateParams& p) {
    // constructor (dsp/deesser.rs:110-134): bounds 4000 / 11000 split into thirds (:242-255)
    auto bounds = [](double lo, double hi, double out[4]) {
        const double span = rmax(hi - lo, 600.0);
        out[0] = lo;
        out[1] = lo + span / 3.0;
        out[2] = lo + span * 2.0 / 3.0;
        out[3] = hi;
    };
    auto center = [](double lo, double hi) { return std::sqrt(lo * hi); };                                   // :258-260
    auto dyn_q = [&](double lo, double hi) { return rclamp(center(lo, hi) / rmax(hi - lo, 200.0), 0.5, 6.0); };  // :263-266
    double e0[4];
    bounds(4000.0, 11000.0, e0);
    // setters in python_api.rs:420-436 order; each clamps (dsp/deesser.rs:303-353)
    double low = 4000.0, high = 11000.0;
    low = rclamp(s.deesser_low_cut_hz, 2000.0, 12000.0);
    if (high <= low + 200.0) high = rclamp(low + 200.0, 2200.0, 16000.0);
    high = rclamp(s.deesser_high_cut_hz, 2200.0, 16000.0);
    if (high <= low + 200.0) low = rclamp(high - 200.0, 2000.0, 12000.0);
    double e1[4];
    bounds(low, high, e1);
    for (int b = 0; b < 3; ++b) {
        store(p.de_det0[2 * b], design_biquad(BqKind::HighPass, e0[b], 0.0, 0.707, fs));
        store(p.de_det0[2 * b + 1], design_biquad(BqKind::LowPass, e0[b + 1], 0.0, 0.707, fs));
        store(p.de_dyn0[b], design_biquad(BqKind::Peaking, center(e0[b], e0[b + 1]), 0.0, dyn_q(e0[b], e0[b + 1]), fs));
        store(&p.de[DE_DET + 10 * b], design_biquad(BqKind::HighPass, e1[b], 0.0, 0.707, fs));
        store(&p.de[DE_DET + 10 * b + 5], design_biquad(BqKind::LowPass, e1[b + 1], 0.0, 0.707, fs));
        const double f = center(e1[b], e1[b + 1]);
        const double q = rmax(dyn_q(e1[b], e1[b + 1]), 1e-6);
        store(p.de_dyn1[b], design_biquad(BqKind::Peaking, f, 0.0, q, fs));
        const double omega = 2.0 * kPi * f / fs;  // the pieces design() rebuilds when the gain moves
        p.de[DE_DYN_COS + b] = std::cos(omega);
        p.de[DE_DYN_ALPHA + b] = std::sin(omega) / (2.0 * q);
    }
    const double amount = rclamp(s.deesser_auto_amount, 0.0, 1.0);
    const double threshold = rclamp(s.deesser_threshold_db, -60.0, -6.0);
    const double ratio = rclamp(s.deesser_ratio, 1.0, 20.0);
    const double max_red = rclamp(s.deesser_max_reduction_db, 0.0, 24.0);
    p.de[DE_ATTACK] = time_constant_to_coeff(rclamp(s.deesser_attack_ms, 0.1, 50.0), fs);
    p.de[DE_RELEASE] = time_constant_to_coeff(rclamp(s.deesser_release_ms, 5.0, 500.0), fs);
    p.de[DE_DET_ATTACK] = time_constant_to_coeff(1.5, fs);   // :123-124
    p.de[DE_DET_RELEASE] = time_constant_to_coeff(60.0, fs);
    p.de[DE_MAX_RED] = max_red;
    p.de[DE_THRESHOLD] = threshold;
    p.de[DE_RATIO_FACTOR] = 1.0 - (1.0 / ratio);
    p.de[DE_RATIO_THR] = rclamp((threshold + 60.0) * 0.10, 0.0, 6.0);
    p.de[DE_TRIGGER] = lerp(8.0, 0.8, amount);               // :446-450
    p.de[DE_SLOPE] = lerp(0.08, 1.9, amount);
    p.de[DE_CAP] = rmin(lerp(0.8, 14.0, amount), max_red * 0.75);
    p.de[DE_CONF_FLOOR] = rclamp(lerp(0.28, 0.06, amount), 0.0, 0.95);
    p.de[DE_BASE_FALL] = time_constant_to_coeff(13.88, fs);
    p.de[DE_BASE_RISE] = time_constant_to_coeff(34.72, fs);
    p.de[DE_BASE_INACTIVE] = time_constant_to_coeff(20.82, fs);
    p.de[DE_MANUAL_CAP] = max_red * 0.75;
    if (s.deesser_auto_enabled) p.flags |= LF_DE_AUTO;
    return e1[1] >= 0.5 * fs || e1[2] >= 0.5 * fs || e1[3] >= 0.5 * fs;  // a configured edge at / beyond Nyquist
}

void plan_compressor(const AfChainSettings& s, double fs, CandidateParams& p) {
    // OfflineDspBlockProcessor::new builds -18 dB / 3:1 / 5 ms / 100 ms / 0 dB / knee 6
    // (block_processor.rs:50); python_api.rs:440-466 then calls the setters in a fixed order.
    const bool adaptive = s.compressor_adaptive_release != 0;
    p.c_threshold = s.compressor_threshold_db;
    p.c_factor = 1.0 - 1.0 / rmax(s.compressor_ratio, 1.0);
    p.c_knee = 6.0;
    p.c_attack = time_constant_to_coeff(s.compressor_attack_ms, fs);
    p.c_det_release = time_constant_to_coeff(s.compressor_release_ms, fs);  // dsp/compressor.rs:236-244
    // Gain-reduction release when not adaptive: set_base_release_time runs last and overwrites it
    // with base_release_ms (dsp/compressor.rs:268-275); adaptive mode never reads it (:468-505).
    p.c_release = time_constant_to_coeff(s.compressor_base_release_ms, fs);
    p.c_rms = time_constant_to_coeff(20.0, fs);
    p.c_makeup_lin = db_to_linear(s.compressor_makeup_gain_db);  // smoothed makeup == manual makeup (:294-300)
    {
        const double cutoff = rclamp(120.0, 20.0, fs * 0.45);     // dsp/compressor.rs:390-394
        const double omega = 2.0 * kPi * cutoff / rmax(fs, 1.0);
        p.c_sc = 1.0 / (1.0 + omega);
    }
    p.c_band = time_constant_to_coeff(18.0, fs);
    p.c_fast = time_constant_to_coeff(50.0, fs);
    p.c_charge = time_constant_to_coeff(250.0, fs);
    p.c_slow = time_constant_to_coeff(400.0, fs);
    if (adaptive) p.flags |= LF_C_ADAPTIVE;
    if (s.compressor_sidechain_highpass_enabled) p.flags |= LF_C_SIDECHAIN;
    // auto makeup (dsp/compressor.rs:318-331,598-653): set_makeup_gain runs before set_auto_makeup_enabled, so the
    // smoothed makeup starts at the manual value; without a meter for the rate the switch stays off (:319)
    p.c_makeup_db = s.compressor_makeup_gain_db;
    p.c_target_lufs = rclamp(s.compressor_target_lufs, -24.0, -12.0);
    p.c_mk_smooth = time_constant_to_coeff(200.0, fs);
    p.c_mk_activity = time_constant_to_coeff(200.0, fs);
    p.c_mk_relax = time_constant_to_coeff(1500.0, fs);
    if (s.compressor_auto_makeup_enabled && loudness_meter_supports(fs)) p.flags |= LF_C_AUTO_MAKEUP;
}

int gcd_int(int a, int b) {
    while (b) {
        const int t = a % b;
        a = b;
        b = t;
    }
    return a;
}

}  // namespace

double time_constant_to_coeff(double time_ms, double sample_rate) {  // dsp/util.rs:5-9
    const double tau = rmax(time_ms, 0.001) / 1000.0;
    return std::exp(-1.0 / (tau * sample_rate));
}

// dsp/biquad.rs:110-182 (RBJ cookbook forms, normalised by a0)
BiquadCoeffs design_biquad(BqKind kind, double frequency_hz, double gain_db, double q_in, double sample_rate) {
    const double omega = 2.0 * kPi * frequency_hz / sample_rate;
    const double sn = std::sin(omega);
    const double cs = std::cos(omega);
    const double q = rmax(q_in, 1e-6);
    const double alpha = sn / (2.0 * q);
    const double a = std::pow(10.0, gain_db / 40.0);
    double b0, b1, b2, a0, a1, a2;
    switch (kind) {
        case BqKind::Peaking:
            b0 = 1.0 + alpha * a;
            b1 = -2.0 * cs;
            b2 = 1.0 - alpha * a;
            a0 = 1.0 + alpha / a;
            a1 = -2.0 * cs;
            a2 = 1.0 - alpha / a;
            break;
        case BqKind::LowShelf: {
            const double t = 2.0 * std::sqrt(a) * alpha;
            b0 = a * ((a + 1.0) - (a - 1.0) * cs + t);
            b1 = 2.0 * a * ((a - 1.0) - (a + 1.0) * cs);
            b2 = a * ((a + 1.0) - (a - 1.0) * cs - t);
            a0 = (a + 1.0) + (a - 1.0) * cs + t;
            a1 = -2.0 * ((a - 1.0) + (a + 1.0) * cs);
            a2 = (a + 1.0) + (a - 1.0) * cs - t;
            break;
        }
        case BqKind::HighShelf: {
            const double t = 2.0 * std::sqrt(a) * alpha;
            b0 = a * ((a + 1.0) + (a - 1.0) * cs + t);
            b1 = -2.0 * a * ((a - 1.0) + (a + 1.0) * cs);
            b2 = a * ((a + 1.0) + (a - 1.0) * cs - t);
            a0 = (a + 1.0) - (a - 1.0) * cs + t;
            a1 = 2.0 * ((a - 1.0) - (a + 1.0) * cs);
            a2 = (a + 1.0) - (a - 1.0) * cs - t;
            break;
        }
        case BqKind::Notch:
            b0 = 1.0;
            b1 = -2.0 * cs;
            b2 = 1.0;
            a0 = 1.0 + alpha;
            a1 = -2.0 * cs;
            a2 = 1.0 - alpha;
            break;
        case BqKind::HighPass:
            b0 = (1.0 + cs) / 2.0;
            b1 = -(1.0 + cs);
            b2 = (1.0 + cs) / 2.0;
            a0 = 1.0 + alpha;
            a1 = -2.0 * cs;
            a2 = 1.0 - alpha;
            break;
        default:  // LowPass
            b0 = (1.0 - cs) / 2.0;
            b1 = 1.0 - cs;
            b2 = (1.0 - cs) / 2.0;
            a0 = 1.0 + alpha;
            a1 = -2.0 * cs;
            a2 = 1.0 - alpha;
            break;
    }
    return {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
}

RateConstants rate_constants(double fs) {
    RateConstants rc;
    rc.block_samples = static_cast<int>(std::clamp<size_t>(as_usize(std::round(fs * 0.020)), 1, 8192));
    const double fade = std::round(fs * 1.5 / 1000.0);  // dsp/biquad.rs:12-19
    rc.fade_samples = std::isfinite(fade) ? static_cast<int>(std::clamp<size_t>(as_usize(fade), 1, 4096)) : 1;
    for (size_t i = 0; i < 10; ++i)
        store(rc.eq_default[i], design_biquad(default_kind(i), kDefaultFrequencies[i], 0.0, kDefaultQ, fs));
    // adaptive input cleanup (audio/processor/routing.rs): everything below is f32 like the reference
    CleanupConst& k = rc.cleanup;
    const float pi = 3.14159265358979323846f;
    const float fsf = static_cast<float>(fs);
    k.fs = fsf;
    k.lowpass_coeff = rclampf(2.0f * pi * 150.0f / fsf, 0.0f, 1.0f);                 // routing.rs:341
    for (int i = 0; i < kHumBins; ++i) {
        for (int h = 0; h < 2; ++h) {                                                // HumBin::new, routing.rs:64-76
            const float f = (h == 0 ? 1.0f : 2.0f) * (49.0f + static_cast<float>(i) * 1.0f);
            const float omega = 2.0f * pi * f / std::fmax(fsf, 1.0f);
            k.bin_cos[h * kHumBins + i] = std::cos(omega);
            k.bin_sin[h * kHumBins + i] = std::sin(omega);
        }
    }
    for (int h = 0; h < 2; ++h) {                                                    // NotchFilter::new, routing.rs:117-140
        const float f = h == 0 ? 55.0f : 110.0f;
        const float omega = 2.0f * pi * f / std::fmax(fsf, 1.0f);
        const float sn = std::sin(omega), cs = std::cos(omega);
        const float alpha = sn / (2.0f * std::fmax(36.0f, 1.0f));
        const float a0 = 1.0f + alpha;
        k.notch0[h][0] = 1.0f / a0;
        k.notch0[h][1] = -2.0f * cs / a0;
        k.notch0[h][2] = 1.0f / a0;
        k.notch0[h][3] = -2.0f * cs / a0;
        k.notch0[h][4] = (1.0f - alpha) / a0;
    }
    const double hp_hz[3] = {80.0, 100.0, 120.0};                                    // routing.rs:539-547
    for (int i = 0; i < 3; ++i) store(k.hp[i], design_biquad(BqKind::HighPass, hp_hz[i], 0.0, 0.707, static_cast<double>(fsf)));
    auto as_u32 = [](float v) -> uint32_t {
        if (!(v > 0.0f)) return 0;
        if (v >= 4294967296.0f) return 0xffffffffu;
        return static_cast<uint32_t>(v);
    };
    k.window_samples = static_cast<int>(as_usize(static_cast<double>(std::fmax(std::round(fsf * 0.25f), 1.0f))));
    k.notch_fade_total = static_cast<int>(as_usize(static_cast<double>(std::fmax(std::round(fsf * 0.020f), 1.0f))));
    k.hp_fade_total = rc.fade_samples;
    k.rumble_hold_gentle = as_u32(std::round(fsf * 0.18f));
    k.rumble_hold_strong = as_u32(std::round(fsf * 0.30f));
    k.hum_hold = as_u32(std::round(fsf * 0.75f));
    return rc;
}

int validate_typed_bands(const AfBand bands[AFSIM_NUM_BANDS], double fs, std::string* error) {  // lib.rs:154-189
    if (!std::isfinite(fs) || fs <= 0.0) return fail(error, AFSIM_INVALID_ARGUMENT, "sample_rate must be finite and positive");
    for (size_t i = 0; i < AFSIM_NUM_BANDS; ++i) {
        if (bands[i].filter_type > AF_LOW_PASS)
            return fail(error, AFSIM_INVALID_ARGUMENT,
                        "band " + std::to_string(i) + " has unsupported EQ filter type: " + std::to_string(bands[i].filter_type));
        const std::string msg = validate_band(bands[i], i, fs);
        if (!msg.empty()) return fail(error, AFSIM_INVALID_ARGUMENT, msg);
    }
    return AFSIM_OK;
}

int validate_legacy_response_bands(const AfBand bands[AFSIM_NUM_BANDS], double fs, std::string* error) {  // lib.rs:105-133
    if (!std::isfinite(fs) || fs <= 0.0) return fail(error, AFSIM_INVALID_ARGUMENT, "sample_rate must be finite and positive");
    const double nyquist = fs / 2.0;
    for (size_t i = 0; i < AFSIM_NUM_BANDS; ++i) {
        const AfBand& b = bands[i];
        const std::string idx = std::to_string(i);
        if (!std::isfinite(b.frequency_hz) || b.frequency_hz <= 0.0 || b.frequency_hz >= nyquist)
            return fail(error, AFSIM_INVALID_ARGUMENT, "band " + idx + " frequency must be between 0 Hz and Nyquist");
        if (!std::isfinite(b.gain_db)) return fail(error, AFSIM_INVALID_ARGUMENT, "band " + idx + " gain must be finite");
        if (!std::isfinite(b.q) || b.q <= 0.0)
            return fail(error, AFSIM_INVALID_ARGUMENT, "band " + idx + " Q must be finite and positive");
    }
    return AFSIM_OK;
}

int validate_response_frequencies(const double* freqs, size_t n, double fs, std::string* error) {  // lib.rs:134-141,199-206
    const double nyquist = fs / 2.0;
    for (size_t i = 0; i < n; ++i)
        if (!std::isfinite(freqs[i]) || freqs[i] < 0.0 || freqs[i] > nyquist)
            return fail(error, AFSIM_INVALID_ARGUMENT, "response frequencies must be finite and between 0 Hz and Nyquist");
    return AFSIM_OK;
}

void plan_eq_sections(const AfBand bands[AFSIM_NUM_BANDS], bool typed, double fs, double coeffs[kMaxSections][5],
                      int band_sections[AFSIM_NUM_BANDS]) {
    for (int i = 0; i < kMaxSections; ++i) store_identity(coeffs[i]);
    design_eq(bands, typed, fs, coeffs, /*band_major_slots=*/true, band_sections);
}

int plan_candidate(const AfBand bands[AFSIM_NUM_BANDS], const AfChainSettings& s, double fs, CandidatePlan* out,
                   std::string* error) {
    if (!std::isfinite(fs) || fs <= 0.0)  // python_api.rs:388-392
        return fail(error, AFSIM_INVALID_ARGUMENT, "sample_rate must be positive and finite");
    if (s.use_typed_bands) {
        const int rc = validate_typed_bands(bands, fs, error);
        if (rc != AFSIM_OK) return rc;
    }
    if (s.input_stage > AF_INPUT_CLEANUP_STRONG) return fail(error, AFSIM_INVALID_ARGUMENT, "unknown input_stage");

    std::memset(out, 0, sizeof *out);
    CandidateParams& p = out->params;
    for (int i = 0; i < kMaxSections; ++i) store_identity(p.eq[i]);
    p.n_sections = static_cast<uint32_t>(design_eq(bands, s.use_typed_bands != 0, fs, p.eq, /*band_major_slots=*/false, nullptr));
    if (!s.use_typed_bands) p.flags |= LF_EQ_FADE;  // python_api.rs:407-413: setters without reset()

    out->structure = ST_EQ;
    if (s.eq_before_deesser) out->structure |= ST_EQ_BEFORE_DEESSER;
    if (s.deesser_enabled) {
        out->structure |= ST_DEESSER;
        out->deesser_unstable = plan_deesser(s, fs, p) ? 1u : 0u;
    }
    if (s.compressor_enabled) {
        out->structure |= ST_COMPRESSOR;
        plan_compressor(s, fs, p);
        if (p.flags & LF_C_AUTO_MAKEUP) out->structure |= ST_AUTO_MAKEUP;
    }
    // python_api.rs:470-487
    const float effective_ceiling =
        static_cast<float>(effective_limiter_ceiling_db(s.limiter_ceiling_db, s.limiter_careful_output_enabled != 0));
    p.effective_ceiling_db = effective_ceiling;
    p.tp_ceil = 1.0f;
    p.tp_release = 0.0f;
    p.l_ceil = 1.0;
    if (s.limiter_enabled) {
        out->structure |= ST_LIMITER;
        const double lookahead = std::round(rclamp(s.limiter_lookahead_ms, 0.1, 10.0) / 1000.0 * fs);  // dsp/limiter.rs:113-115
        out->lookahead = static_cast<uint32_t>(std::clamp<size_t>(as_usize(lookahead), 1, kMaxLookahead));
        const double ceiling_db = rmin(static_cast<double>(effective_ceiling), 0.0);                    // dsp/limiter.rs:139-142
        p.l_ceil = db_to_linear(ceiling_db);
        p.l_release = time_constant_to_coeff(s.limiter_release_ms, fs);
        // block_processor.rs:150-151 + dsp/true_peak.rs:304-313 (f32 arithmetic)
        p.tp_ceil = rclampf(std::pow(10.0f, static_cast<float>(ceiling_db) / 20.0f), 0.000001f, 1.0f);
        const float tp_fs = std::fmax(static_cast<float>(fs), 1.0f);
        const float release_ms = rclampf(static_cast<float>(s.limiter_release_ms), 5.0f, 500.0f);
        p.tp_release = static_cast<float>(time_constant_to_coeff(static_cast<double>(release_ms), static_cast<double>(tp_fs)));
    }
    out->input_stage = s.input_stage;
    if (s.input_stage == AF_INPUT_DC_HP80)  // processor.rs:74-76
        store(p.in_hp, design_biquad(BqKind::HighPass, 80.0, 0.0, 0.707, fs));
    return AFSIM_OK;
}

// Loudness meter constants (dsp/loudness.rs:99-131 over the `ebur128` crate 0.1.10, mode M).  The crate is not in the
// reference tree; it ports libebur128, whose K-weighting is the BS.1770 high shelf (f0 1681.97 Hz, +4 dB, Q 0.7072)
// times the RLB high-pass (f0 38.135 Hz, Q 0.5003), re-derived for the sample rate and folded into one 4th-order
// section -- parity with the crate is unpinned.
MakeupConst makeup_constants(double fs, int block_samples, int n_samples) {
    MakeupConst m;
    std::memset(&m, 0, sizeof m);
    const double rate = static_cast<double>(static_cast<uint32_t>(std::min<size_t>(as_usize(fs), 0xffffffffu)));
    const double f0 = 1681.974450955533, G = 3.999843853973347, Q = 0.7071752369554196;
    double K = std::tan(kPi * f0 / rate);
    const double Vh = std::pow(10.0, G / 20.0);
    const double Vb = std::pow(Vh, 0.4996667741545416);
    const double a0 = 1.0 + K / Q + K * K;
    const double pb[3] = {(Vh + Vb * K / Q + K * K) / a0, 2.0 * (K * K - Vh) / a0, (Vh - Vb * K / Q + K * K) / a0};
    const double pa[3] = {1.0, 2.0 * (K * K - 1.0) / a0, (1.0 - K / Q + K * K) / a0};
    const double f1 = 38.13547087602444, Q1 = 0.5003270373238773;
    K = std::tan(kPi * f1 / rate);
    const double rb[3] = {1.0, -2.0, 1.0};
    const double ra[3] = {1.0, 2.0 * (K * K - 1.0) / (1.0 + K / Q1 + K * K), (1.0 - K / Q1 + K * K) / (1.0 + K / Q1 + K * K)};
    m.b[0] = pb[0] * rb[0];
    m.b[1] = pb[0] * rb[1] + pb[1] * rb[0];
    m.b[2] = pb[0] * rb[2] + pb[1] * rb[1] + pb[2] * rb[0];
    m.b[3] = pb[1] * rb[2] + pb[2] * rb[1];
    m.b[4] = pb[2] * rb[2];
    m.a[0] = pa[0] * ra[0];
    m.a[1] = pa[0] * ra[1] + pa[1] * ra[0];
    m.a[2] = pa[0] * ra[2] + pa[1] * ra[1] + pa[2] * ra[0];
    m.a[3] = pa[1] * ra[2] + pa[2] * ra[1];
    m.a[4] = pa[2] * ra[2];
    m.window = static_cast<int>(static_cast<uint64_t>(rate) * 400 / 1000);
    const int block = std::max(block_samples, 1);
    m.slot = std::max(gcd_int(block, std::max(m.window, 1)), 1);
    m.n_slots = std::max(m.window / m.slot, 1);
    m.tail_from = (n_samples % block) % m.slot;
    return m;
}

// simulate_auto_makeup_control (python_api.rs:118-276): a compressor alone, knee 6, auto makeup forced on.
int plan_makeup_control(const AfAutoMakeupSettings& s, double fs, double noise_floor_db, double noise_reliability,
                        bool has_vad, CandidatePlan* out, std::string* error) {
    if (!std::isfinite(fs) || fs <= 0.0) return fail(error, AFSIM_INVALID_ARGUMENT, "sample_rate must be positive and finite");
    if (!std::isfinite(noise_floor_db) || !std::isfinite(noise_reliability) || !(noise_reliability >= 0.0 && noise_reliability <= 1.0))
        return fail(error, AFSIM_INVALID_ARGUMENT, "noise evidence must be finite and reliability must be between 0 and 1");
    if (!std::isfinite(s.vad_reliability) || !(s.vad_reliability >= 0.0 && s.vad_reliability <= 1.0))
        return fail(error, AFSIM_INVALID_ARGUMENT, "vad_reliability must be finite and between 0 and 1");
    std::memset(out, 0, sizeof *out);
    CandidateParams& p = out->params;
    for (int i = 0; i < kMaxSections; ++i) store_identity(p.eq[i]);
    AfChainSettings cs;
    chain_settings_default(&cs);
    cs.compressor_threshold_db = s.threshold_db;
    cs.compressor_ratio = s.ratio;
    cs.compressor_attack_ms = s.attack_ms;
    cs.compressor_release_ms = s.release_ms;
    // Compressor::new keeps release_ms as the base release (dsp/compressor.rs:133-202); set_adaptive_release(false)
    // re-derives the gain-reduction release from it (:247-260)
    cs.compressor_base_release_ms = s.release_ms;
    cs.compressor_makeup_gain_db = s.makeup_gain_db;
    cs.compressor_target_lufs = s.target_lufs;
    cs.compressor_adaptive_release = s.adaptive_release;
    cs.compressor_sidechain_highpass_enabled = s.sidechain_highpass_enabled;
    cs.compressor_auto_makeup_enabled = 1;
    plan_compressor(cs, fs, p);
    out->structure = ST_COMPRESSOR;
    if (p.flags & LF_C_AUTO_MAKEUP) out->structure |= ST_AUTO_MAKEUP;  // no meter for this rate: the switch stays off
    p.c_ev_vad_reliability = s.vad_reliability;
    p.c_ev_noise_floor_db = noise_floor_db;
    p.c_ev_live_reliability = noise_reliability;
    p.c_ev_cfg_reliability = noise_reliability;  // set_noise_reference_reliability (python_api.rs:172)
    if (has_vad) p.flags |= LF_C_EVIDENCE;
    p.tp_ceil = 1.0f;
    p.l_ceil = 1.0;
    p.effective_ceiling_db = 0.0f;
    return AFSIM_OK;
}

int plan_eq_only(const AfBand bands[AFSIM_NUM_BANDS], double fs, CandidatePlan* out, std::string* error) {
    const int rc = validate_typed_bands(bands, fs, error);
    if (rc != AFSIM_OK) return rc;
    std::memset(out, 0, sizeof *out);
    CandidateParams& p = out->params;
    for (int i = 0; i < kMaxSections; ++i) store_identity(p.eq[i]);
    p.n_sections = static_cast<uint32_t>(design_eq(bands, true, fs, p.eq, false, nullptr));
    p.tp_ceil = 1.0f;
    p.l_ceil = 1.0;
    out->structure = ST_EQ | ST_INPUT_TRUE_PEAK;
    return AFSIM_OK;
}

void chain_settings_default(AfChainSettings* s) {  // python_api.rs:415-487 defaults
    std::memset(s, 0, sizeof *s);
    s->deesser_auto_enabled = 1;
    s->compressor_enabled = 1;
    s->compressor_sidechain_highpass_enabled = 1;
    s->limiter_enabled = 1;
    s->limiter_careful_output_enabled = 1;
    s->deesser_auto_amount = 0.5;
    s->deesser_low_cut_hz = 4000.0;
    s->deesser_high_cut_hz = 11000.0;
    s->deesser_threshold_db = -28.0;
    s->deesser_ratio = 4.0;
    s->deesser_attack_ms = 2.0;
    s->deesser_release_ms = 80.0;
    s->deesser_max_reduction_db = 6.0;
    s->compressor_threshold_db = -20.0;
    s->compressor_ratio = 4.0;
    s->compressor_attack_ms = 10.0;
    s->compressor_release_ms = 200.0;
    s->compressor_makeup_gain_db = 0.0;
    s->compressor_base_release_ms = 50.0;
    s->compressor_target_lufs = -18.0;
    s->limiter_ceiling_db = -0.5;
    s->limiter_release_ms = 50.0;
    s->limiter_lookahead_ms = 2.0;
}

void default_bands(AfBand out[AFSIM_NUM_BANDS]) {  // dsp/eq.rs:11-23,125-140
    for (size_t i = 0; i < AFSIM_NUM_BANDS; ++i) {
        std::memset(&out[i], 0, sizeof out[i]);
        out[i].frequency_hz = kDefaultFrequencies[i];
        out[i].gain_db = 0.0;
        out[i].q = kDefaultQ;
        out[i].filter_type = i == 0 ? AF_LOW_SHELF : (i == 9 ? AF_HIGH_SHELF : AF_BELL);
        out[i].slope_db_per_octave = 12;
        out[i].enabled = 1;
    }
}

std::vector<std::vector<uint32_t>> cut_stream_group(const std::vector<uint32_t>& passage, const std::vector<uint32_t>& eq_class,
                                                    int max_streams) {
    const size_t S = passage.size();
    std::vector<uint32_t> order(S);
    for (size_t i = 0; i < S; ++i) order[i] = static_cast<uint32_t>(i);
    std::vector<std::vector<uint32_t>> pieces;
    bool cut = max_streams >= 32 && S > static_cast<size_t>(max_streams) && eq_class.size() == S;
    if (cut) {
        std::set<std::pair<uint32_t, uint32_t>> distinct;
        for (size_t i = 0; i < S && distinct.size() * 4 <= static_cast<size_t>(max_streams); ++i)
            distinct.emplace(passage[i], eq_class[i]);
        cut = distinct.size() * 4 <= static_cast<size_t>(max_streams);
    }
    if (!cut) {
        pieces.push_back(std::move(order));
        return pieces;
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return passage[x] < passage[y]; });
    const size_t n_pieces = (S + max_streams - 1) / max_streams;
    const size_t piece_len = ((S + n_pieces - 1) / n_pieces + 31) / 32 * 32;  // <= max_streams when that is a multiple of 32
    const size_t len = std::min(piece_len, static_cast<size_t>(max_streams));
    for (size_t first = 0; first < S; first += len)
        pieces.emplace_back(order.begin() + first, order.begin() + std::min(S, first + len));
    return pieces;
}

}  // namespace afsim
