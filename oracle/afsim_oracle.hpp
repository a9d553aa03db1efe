// afsim_oracle.hpp -- CPU restatement of AudioForge's offline chain simulator.
//
// TEST INFRASTRUCTURE ONLY.  This is the parity oracle: a scalar, single-threaded
// restatement of the reference's Rust DSP, function by function, in the
// reference's own evaluation order (f64 state / f32 hand-off between stages, no
// FMA contraction except the explicit fused FIR).  Only tests/, the smoke check
// in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
// build, load or call it.  The product (audio_forge_b200/csrc) never does.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
//
// Parity pin: reproduces the reference's golden vector
// rust-core/src/audio/processor/tests.rs:1784-1885 and the known-answer tests
// listed in DESIGN.md (tests/test_oracle_*.py), and three evaluation reports the
// real reference published from its native core (evaluation/processing-order-,
// dynamics-aliasing-, limiter-lookahead-report.json: de-esser / EQ order over the
// generated 96-clip corpus, compressor at 48 / 192 kHz, limiter lookaheads) when
// the reference's own tool code runs with this oracle as its native core -- bit
// for bit but for two f32 log10 conversions one ulp apart
// (tests/test_oracle_reference_report.py).  The auto-makeup loudness meter
// restates the third-party `ebur128` crate 0.1.10 (absent from the reference
// tree): that sub-path is "parity unpinned" at source level; nine of the eleven
// numbers of the reference's published auto-makeup benchmark are reproduced to nine
// decimals (tests/test_oracle_auto_makeup_benchmark.py).
//
// Every class cites the reference file:line it follows (paths relative to the
// reference checkout, rust-core/src/...).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace orc {

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr float kPiF = 3.14159265358979323846264338327950288f;

// ---- Rust numeric semantics ---------------------------------------------------
// f64::max/min ignore a NaN operand (== fmax/fmin); clamp keeps NaN.
inline double rmax(double a, double b) { return std::fmax(a, b); }
inline double rmin(double a, double b) { return std::fmin(a, b); }
inline float rmaxf(float a, float b) { return std::fmax(a, b); }
inline float rminf(float a, float b) { return std::fmin(a, b); }
inline double rclamp(double v, double lo, double hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
inline float rclampf(float v, float lo, float hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
// `x as usize` for f64: saturating, NaN -> 0.
inline size_t as_usize(double v) {
    if (!(v > 0.0)) return 0;
    if (v >= 18446744073709551615.0) return std::numeric_limits<size_t>::max();
    return static_cast<size_t>(v);
}
inline uint32_t as_u32(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 4294967296.0f) return 0xffffffffu;
    return static_cast<uint32_t>(v);
}

// ---- dsp/util.rs:5-20 ----------------------------------------------------------
inline double time_constant_to_coeff(double time_ms, double sample_rate) {
    const double tau = rmax(time_ms, 0.001) / 1000.0;
    return std::exp(-1.0 / (tau * sample_rate));
}
inline double db_to_linear(double db) { return std::pow(10.0, db / 20.0); }
inline double linear_to_db(double linear, double min_linear) {
    return 20.0 * std::log10(rmax(std::fabs(linear), min_linear));
}

// ---- dsp/biquad.rs ---------------------------------------------------------------
enum class BiquadType { LowShelf, HighShelf, Peaking, Notch, HighPass, LowPass, Bypass };

struct Coeffs {
    double b0, b1, b2, a1, a2;
};

class Biquad {
  public:
    // biquad.rs:70-107
    Biquad(BiquadType type, double frequency, double gain_db, double q, double sample_rate)
        : type_(type), frequency_(frequency), gain_db_(gain_db), q_(q), sample_rate_(sample_rate) {
        commit(design());
    }

    // biquad.rs:110-182 (RBJ cookbook forms, normalised by a0)
    Coeffs design() const {
        const double omega = 2.0 * kPi * frequency_ / sample_rate_;
        const double sn = std::sin(omega);
        const double cs = std::cos(omega);
        const double q = rmax(q_, 1e-6);
        const double alpha = sn / (2.0 * q);
        const double a = std::pow(10.0, gain_db_ / 40.0);
        double b0, b1, b2, a0, a1, a2;
        switch (type_) {
            case BiquadType::Peaking:
                b0 = 1.0 + alpha * a;
                b1 = -2.0 * cs;
                b2 = 1.0 - alpha * a;
                a0 = 1.0 + alpha / a;
                a1 = -2.0 * cs;
                a2 = 1.0 - alpha / a;
                break;
            case BiquadType::LowShelf: {
                const double t = 2.0 * std::sqrt(a) * alpha;
                b0 = a * ((a + 1.0) - (a - 1.0) * cs + t);
                b1 = 2.0 * a * ((a - 1.0) - (a + 1.0) * cs);
                b2 = a * ((a + 1.0) - (a - 1.0) * cs - t);
                a0 = (a + 1.0) + (a - 1.0) * cs + t;
                a1 = -2.0 * ((a - 1.0) + (a + 1.0) * cs);
                a2 = (a + 1.0) + (a - 1.0) * cs - t;
                break;
            }
            case BiquadType::HighShelf: {
                const double t = 2.0 * std::sqrt(a) * alpha;
                b0 = a * ((a + 1.0) + (a - 1.0) * cs + t);
                b1 = -2.0 * a * ((a - 1.0) + (a + 1.0) * cs);
                b2 = a * ((a + 1.0) + (a - 1.0) * cs - t);
                a0 = (a + 1.0) - (a - 1.0) * cs + t;
                a1 = 2.0 * ((a - 1.0) - (a + 1.0) * cs);
                a2 = (a + 1.0) - (a - 1.0) * cs - t;
                break;
            }
            case BiquadType::Notch:
                b0 = 1.0;
                b1 = -2.0 * cs;
                b2 = 1.0;
                a0 = 1.0 + alpha;
                a1 = -2.0 * cs;
                a2 = 1.0 - alpha;
                break;
            case BiquadType::HighPass:
                b0 = (1.0 + cs) / 2.0;
                b1 = -(1.0 + cs);
                b2 = (1.0 + cs) / 2.0;
                a0 = 1.0 + alpha;
                a1 = -2.0 * cs;
                a2 = 1.0 - alpha;
                break;
            case BiquadType::LowPass:
                b0 = (1.0 - cs) / 2.0;
                b1 = 1.0 - cs;
                b2 = (1.0 - cs) / 2.0;
                a0 = 1.0 + alpha;
                a1 = -2.0 * cs;
                a2 = 1.0 - alpha;
                break;
            default:
                b0 = 1.0; b1 = 0.0; b2 = 0.0; a0 = 1.0; a1 = 0.0; a2 = 0.0;
                break;
        }
        return {b0 / a0, b1 / a0, b2 / a0, a1 / a0, a2 / a0};
    }

    // biquad.rs:184-205
    static double response_db(const Coeffs& c, double frequency_hz, double sample_rate) {
        const double omega = 2.0 * kPi * frequency_hz / sample_rate;
        const double c1 = std::cos(omega), s1 = std::sin(omega);
        const double c2 = std::cos(2.0 * omega), s2 = std::sin(2.0 * omega);
        const double nr = c.b0 + c.b1 * c1 + c.b2 * c2;
        const double ni = -c.b1 * s1 - c.b2 * s2;
        const double dr = 1.0 + c.a1 * c1 + c.a2 * c2;
        const double di = -c.a1 * s1 - c.a2 * s2;
        const double np = nr * nr + ni * ni;
        const double dp = dr * dr + di * di;
        const double mag = std::sqrt(np / rmax(dp, 1.0e-30));
        return 20.0 * std::log10(rmax(mag, 1.0e-10));
    }
    double magnitude_response_db(double f) const {  // biquad.rs:208-217 (active coefficients)
        return enabled_ ? response_db(active_, f, sample_rate_) : 0.0;
    }
    double target_magnitude_response_db(double f) const {  // biquad.rs:220-229 (configured target)
        return enabled_ ? response_db(design(), f, sample_rate_) : 0.0;
    }

    // biquad.rs:277-327: DF2T, optional dual-state crossfade
    float process_sample(float input) {
        if (!enabled_) return input;
        const double x = static_cast<double>(input);
        const double ya = step(x, active_, z1_, z2_);
        if (fade_remaining_ == 0) return static_cast<float>(ya);
        const double yp = step(x, pending_, pz1_, pz2_);
        const size_t pos = fade_total_ - fade_remaining_ + 1;
        const double fade = static_cast<double>(pos) / static_cast<double>(fade_total_);
        const double y = ya * (1.0 - fade) + yp * fade;
        fade_remaining_ -= 1;
        if (fade_remaining_ == 0) {  // promote_pending_coefficients, biquad.rs:276-286
            active_ = pending_;
            z1_ = pz1_;
            z2_ = pz2_;
            fade_total_ = 0;
        }
        return static_cast<float>(y);
    }
    void process_block_inplace(float* buf, size_t n) {
        if (!enabled_) return;
        for (size_t i = 0; i < n; ++i) buf[i] = process_sample(buf[i]);
    }

    void reset() { commit(design()); }  // biquad.rs:341-347
    void set_frequency(double f) { frequency_ = f; schedule(design()); }
    void set_gain_db(double g) { gain_db_ = g; schedule(design()); }
    void set_gain_db_immediate(double g) { gain_db_ = g; commit(design()); }  // biquad.rs:365-368
    void set_q(double q) { q_ = rmax(q, 1e-6); schedule(design()); }
    void set_parameters(BiquadType t, double f, double g, double q) {  // biquad.rs:377-389
        type_ = t; frequency_ = f; gain_db_ = g; q_ = rmax(q, 1e-6);
        schedule(design());
    }
    void set_parameters_immediate(BiquadType t, double f, double g, double q) {  // biquad.rs:395-407
        type_ = t; frequency_ = f; gain_db_ = g; q_ = rmax(q, 1e-6);
        commit(design());
    }
    void set_enabled(bool e) { enabled_ = e; }
    double gain_db() const { return gain_db_; }
    bool is_crossfading() const { return fade_remaining_ != 0; }
    const Coeffs& active() const { return active_; }
    const Coeffs& pending() const { return pending_; }
    size_t fade_remaining() const { return fade_remaining_; }

  private:
    static inline double step(double x, const Coeffs& c, double& z1, double& z2) {  // biquad.rs:262-274
        const double y = c.b0 * x + z1;
        z1 = c.b1 * x - c.a1 * y + z2;
        z2 = c.b2 * x - c.a2 * y;
        return y;
    }
    void commit(const Coeffs& c) {  // set_coefficients_immediate, biquad.rs:231-247 (z1/z2 are kept)
        active_ = c;
        pending_ = c;
        pz1_ = 0.0;
        pz2_ = 0.0;
        fade_total_ = 0;
        fade_remaining_ = 0;
    }
    void schedule(const Coeffs& c) {  // schedule_coefficients_crossfade, biquad.rs:249-260
        pending_ = c;
        pz1_ = z1_;
        pz2_ = z2_;
        const double samples = std::round(sample_rate_ * 1.5 / 1000.0);  // biquad.rs:12-19
        size_t total = 1;
        if (std::isfinite(samples)) total = std::clamp<size_t>(as_usize(samples), 1, 4096);
        fade_total_ = total;
        fade_remaining_ = total;
    }

    Coeffs active_{1, 0, 0, 0, 0}, pending_{1, 0, 0, 0, 0};
    double z1_ = 0, z2_ = 0, pz1_ = 0, pz2_ = 0;
    size_t fade_total_ = 0, fade_remaining_ = 0;
    BiquadType type_;
    double frequency_, gain_db_, q_, sample_rate_;
    bool enabled_ = true;
};

// ---- dsp/eq.rs ---------------------------------------------------------------------
enum class EqFilterType : uint8_t { LowShelf = 0, Bell = 1, HighShelf = 2, Notch = 3, HighPass = 4, LowPass = 5 };

struct EqBandConfig {
    EqFilterType filter_type;
    double frequency_hz, gain_db, q;
    uint8_t slope_db_per_octave;
    bool enabled;
};

constexpr double kDefaultFrequencies[10] = {80.0, 160.0, 320.0, 640.0, 1280.0, 2500.0, 5000.0, 8000.0, 12000.0, 16000.0};
constexpr double kDefaultQ = 1.41;

inline EqBandConfig default_band(size_t index) {  // eq.rs:125-140
    EqFilterType t = index == 0 ? EqFilterType::LowShelf : index == 9 ? EqFilterType::HighShelf : EqFilterType::Bell;
    return {t, kDefaultFrequencies[index], 0.0, kDefaultQ, 12, true};
}
inline bool is_pass(EqFilterType t) { return t == EqFilterType::HighPass || t == EqFilterType::LowPass; }
inline BiquadType to_biquad_type(EqFilterType t) {
    switch (t) {
        case EqFilterType::LowShelf: return BiquadType::LowShelf;
        case EqFilterType::Bell: return BiquadType::Peaking;
        case EqFilterType::HighShelf: return BiquadType::HighShelf;
        case EqFilterType::Notch: return BiquadType::Notch;
        case EqFilterType::HighPass: return BiquadType::HighPass;
        default: return BiquadType::LowPass;
    }
}
inline bool supported_slope(uint8_t s) { return s == 12 || s == 24 || s == 36 || s == 48; }

// eq.rs:154-201 validation; returns "" when valid, else the reference's message.
inline std::string fmt_num(double v) {
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
inline std::string validate_band(const EqBandConfig& c, size_t index, double sample_rate) {
    const std::string prefix = "Band " + std::to_string(index) + ": ";
    if (!std::isfinite(c.frequency_hz)) return prefix + "frequency must be finite";
    if (!std::isfinite(sample_rate) || sample_rate <= 40.0)
        return prefix + "sample rate must be finite and support the EQ frequency range";
    const double max_f = rmax(sample_rate / 2.0 - 1.0, 20.0);
    if (!(c.frequency_hz >= 20.0 && c.frequency_hz <= max_f))
        return prefix + "frequency " + fmt_num(c.frequency_hz) + " Hz out of range [20, " + fmt_num(max_f) + "]";
    if (!std::isfinite(c.gain_db)) return prefix + "gain must be finite";
    if (!(c.gain_db >= -12.0 && c.gain_db <= 12.0))
        return prefix + "gain " + fmt_num(c.gain_db) + " dB out of range [-12, 12]";
    if (!std::isfinite(c.q)) return prefix + "Q must be finite";
    if (!(c.q >= 0.1 && c.q <= 10.0)) return prefix + "Q " + fmt_num(c.q) + " out of range [0.1, 10]";
    if (!supported_slope(c.slope_db_per_octave))
        return prefix + "slope " + std::to_string(c.slope_db_per_octave) +
               " dB/octave is unsupported; expected one of [12, 24, 36, 48]";
    return "";
}

inline double butterworth_section_q(size_t section_index, size_t section_count) {  // eq.rs:203-207
    const size_t order = 2 * section_count;
    const double angle = static_cast<double>(2 * section_index + 1) * kPi / static_cast<double>(2 * order);
    return 1.0 / (2.0 * std::cos(angle));
}

class EqBand {  // eq.rs:215-350
  public:
    EqBand(const EqBandConfig& config, double sample_rate) : config_(config) {
        target_sections_ = required_sections(config);
        for (size_t s = 0; s < 4; ++s) {
            if (s < target_sections_) {
                auto [t, g, q] = section_parameters(config, s, target_sections_);
                sections_.emplace_back(t, config.frequency_hz, g, q, sample_rate);
            } else {
                sections_.emplace_back(BiquadType::Bypass, config.frequency_hz, 0.0, kDefaultQ, sample_rate);
            }
        }
        processing_sections_ = target_sections_;
    }
    static size_t required_sections(const EqBandConfig& c) {
        if (!c.enabled) return 0;
        if (is_pass(c.filter_type)) return supported_slope(c.slope_db_per_octave) ? c.slope_db_per_octave / 12 : 1;
        return 1;
    }
    struct SectionParams { BiquadType type; double gain_db; double q; };
    static SectionParams section_parameters(const EqBandConfig& c, size_t index, size_t count) {
        if (is_pass(c.filter_type)) return {to_biquad_type(c.filter_type), 0.0, butterworth_section_q(index, count)};
        const double g = c.filter_type == EqFilterType::Notch ? 0.0 : c.gain_db;
        return {to_biquad_type(c.filter_type), g, c.q};
    }
    void set_config(const EqBandConfig& c) {  // eq.rs:279-298
        config_ = c;
        const size_t target = required_sections(c);
        const size_t processing = std::max(processing_sections_, target);
        for (size_t s = 0; s < processing; ++s) {
            if (s < target) {
                auto [t, g, q] = section_parameters(c, s, target);
                sections_[s].set_parameters(t, c.frequency_hz, g, q);
            } else {
                sections_[s].set_parameters(BiquadType::Bypass, c.frequency_hz, 0.0, kDefaultQ);
            }
        }
        processing_sections_ = processing;
        target_sections_ = target;
    }
    void finish_retired_sections() {  // eq.rs:300-306
        while (processing_sections_ > target_sections_ && !sections_[processing_sections_ - 1].is_crossfading())
            processing_sections_ -= 1;
    }
    float process_sample(float x) {
        for (size_t s = 0; s < processing_sections_; ++s) x = sections_[s].process_sample(x);
        finish_retired_sections();
        return x;
    }
    void process_block_inplace(float* buf, size_t n) {  // eq.rs:317-322 (section-major)
        for (size_t s = 0; s < processing_sections_; ++s) sections_[s].process_block_inplace(buf, n);
        finish_retired_sections();
    }
    void reset() {  // eq.rs:324-336
        const size_t target = required_sections(config_);
        for (size_t s = 0; s < 4; ++s) {
            if (s < target) {
                auto [t, g, q] = section_parameters(config_, s, target);
                sections_[s].set_parameters_immediate(t, config_.frequency_hz, g, q);
            } else {
                sections_[s].set_parameters_immediate(BiquadType::Bypass, config_.frequency_hz, 0.0, kDefaultQ);
            }
        }
        processing_sections_ = target;
        target_sections_ = target;
    }
    double target_magnitude_response_db(double f) const {  // eq.rs:338-343
        double sum = 0.0;
        for (size_t s = 0; s < target_sections_; ++s) sum += sections_[s].target_magnitude_response_db(f);
        return sum;
    }
    const EqBandConfig& config() const { return config_; }
    size_t processing_sections() const { return processing_sections_; }
    const Biquad& section(size_t s) const { return sections_[s]; }

  private:
    std::vector<Biquad> sections_;
    EqBandConfig config_;
    size_t processing_sections_ = 0, target_sections_ = 0;
};

class ParametricEQ {  // eq.rs:352-528
  public:
    explicit ParametricEQ(double sample_rate) : sample_rate_(sample_rate) {
        for (size_t i = 0; i < 10; ++i) bands_.emplace_back(default_band(i), sample_rate);
    }
    void process_block_inplace(float* buf, size_t n) {
        if (!enabled_) return;
        for (auto& b : bands_) b.process_block_inplace(buf, n);
    }
    float process_sample(float x) {
        if (!enabled_) return x;
        for (auto& b : bands_) x = b.process_sample(x);
        return x;
    }
    void reset() { for (auto& b : bands_) b.reset(); }
    void set_band_gain(size_t i, double g) { if (i < 10) { auto c = bands_[i].config(); c.gain_db = g; bands_[i].set_config(c); } }
    void set_band_frequency(size_t i, double f) { if (i < 10) { auto c = bands_[i].config(); c.frequency_hz = f; bands_[i].set_config(c); } }
    void set_band_q(size_t i, double q) { if (i < 10) { auto c = bands_[i].config(); c.q = q; bands_[i].set_config(c); } }
    void set_band_config(size_t i, const EqBandConfig& c) { if (i < 10) bands_[i].set_config(c); }
    void set_enabled(bool e) { enabled_ = e; }
    bool is_enabled() const { return enabled_; }
    std::vector<double> magnitude_response_db(const double* freqs, size_t n) const {  // eq.rs:514-527
        std::vector<double> out(n, 0.0);
        if (!enabled_) return out;
        for (size_t i = 0; i < n; ++i) {
            double sum = 0.0;
            for (const auto& b : bands_) sum += b.target_magnitude_response_db(freqs[i]);
            out[i] = sum;
        }
        return out;
    }
    const EqBand& band(size_t i) const { return bands_[i]; }

  private:
    std::vector<EqBand> bands_;
    bool enabled_ = true;
    double sample_rate_;
};

// ---- dsp/deesser.rs ------------------------------------------------------------------
class DeEsser {
  public:
    explicit DeEsser(double sample_rate) : sample_rate_(sample_rate) {  // deesser.rs:110-134
        attack_coeff_ = time_constant_to_coeff(2.0, sample_rate);
        release_coeff_ = time_constant_to_coeff(80.0, sample_rate);
        detector_attack_coeff_ = time_constant_to_coeff(1.5, sample_rate);
        detector_release_coeff_ = time_constant_to_coeff(60.0, sample_rate);
        const double span = rmax(high_cut_hz_ - low_cut_hz_, 600.0);  // make_bands, deesser.rs:242-255
        const double a = low_cut_hz_ + span / 3.0;
        const double b = low_cut_hz_ + span * 2.0 / 3.0;
        bands_.emplace_back(low_cut_hz_, a, sample_rate);
        bands_.emplace_back(a, b, sample_rate);
        bands_.emplace_back(b, high_cut_hz_, sample_rate);
    }
    static double center_hz(double lo, double hi) { return std::sqrt(lo * hi); }  // deesser.rs:258-260
    static double dynamic_q(double lo, double hi) {                               // deesser.rs:263-266
        const double bw = rmax(hi - lo, 200.0);
        return rclamp(center_hz(lo, hi) / bw, 0.5, 6.0);
    }
    void set_enabled(bool e) { enabled_ = e; }
    bool is_enabled() const { return enabled_; }
    void set_auto_enabled(bool e) { auto_enabled_ = e; }
    void set_auto_amount(double a) { auto_amount_ = rclamp(a, 0.0, 1.0); }
    void set_low_cut_hz(double v) {  // deesser.rs:312-318
        low_cut_hz_ = rclamp(v, 2000.0, 12000.0);
        if (high_cut_hz_ <= low_cut_hz_ + 200.0) high_cut_hz_ = rclamp(low_cut_hz_ + 200.0, 2200.0, 16000.0);
        rebuild();
    }
    void set_high_cut_hz(double v) {  // deesser.rs:321-327
        high_cut_hz_ = rclamp(v, 2200.0, 16000.0);
        if (high_cut_hz_ <= low_cut_hz_ + 200.0) low_cut_hz_ = rclamp(high_cut_hz_ - 200.0, 2000.0, 12000.0);
        rebuild();
    }
    void set_threshold_db(double v) { threshold_db_ = rclamp(v, -60.0, -6.0); }
    void set_ratio(double v) { ratio_ = rclamp(v, 1.0, 20.0); }
    void set_attack_ms(double v) { attack_coeff_ = time_constant_to_coeff(rclamp(v, 0.1, 50.0), sample_rate_); }
    void set_release_ms(double v) { release_coeff_ = time_constant_to_coeff(rclamp(v, 5.0, 500.0), sample_rate_); }
    void set_max_reduction_db(double v) { max_reduction_db_ = rclamp(v, 0.0, 24.0); }
    float current_gain_reduction_db() const { return static_cast<float>(current_reduction_db_); }
    double band_reduction_db(size_t i) const { return bands_[i].reduction_db; }

    float process_sample(float input) {  // deesser.rs:405-547
        if (!enabled_) {
            current_reduction_db_ = 0.0;
            detector_confidence_ = 0.0;
            return input;
        }
        const double broadband_level = static_cast<double>(std::fabs(input));
        broadband_env_ = smooth(broadband_env_, broadband_level, detector_attack_coeff_, detector_release_coeff_);
        double band_level_db[3] = {0, 0, 0};
        double total_env = 0.0, max_env = 0.0;
        for (size_t i = 0; i < 3; ++i) {
            Band& b = bands_[i];
            const float hp = b.detector_hp.process_sample(input);
            const float sc = b.detector_lp.process_sample(hp);
            b.env = smooth(b.env, static_cast<double>(std::fabs(sc)), detector_attack_coeff_, detector_release_coeff_);
            total_env += b.env;
            max_env = rmax(max_env, b.env);
            band_level_db[i] = linear_to_db(b.env, 1e-10);
        }
        const double voice_level = rmax(broadband_env_ - total_env * 0.6, 1e-8);
        const double voice_db = linear_to_db(voice_level, 1e-10);
        const double narrowness = total_env > 1e-10 ? max_env / total_env : 0.0;

        const double amount = rclamp(auto_amount_, 0.0, 1.0);
        const double trigger_offset_db = lerp(8.0, 0.8, amount);
        const double slope = lerp(0.08, 1.9, amount);
        const double auto_cap = lerp(0.8, 14.0, amount);
        const double confidence_floor = lerp(0.28, 0.06, amount);
        const double baseline_fall = time_constant_to_coeff(13.88, sample_rate_);
        const double baseline_rise = time_constant_to_coeff(34.72, sample_rate_);
        const double baseline_inactive = time_constant_to_coeff(20.82, sample_rate_);
        double targets[3] = {0, 0, 0};
        double target_sum = 0.0, aggregate_conf = 0.0;
        for (size_t i = 0; i < 3; ++i) {
            const double level_db = band_level_db[i];
            const double ratio_db = rmax(level_db - voice_db, 0.0);
            const double dominance = max_env > 1e-10 ? std::sqrt(bands_[i].env / max_env) : 0.0;
            const double conf_target = confidence_target(level_db, voice_db, narrowness) * dominance;
            Band& b = bands_[i];
            b.confidence = smooth(b.confidence, rclamp(conf_target, 0.0, 1.0), detector_attack_coeff_, detector_release_coeff_);
            aggregate_conf = rmax(aggregate_conf, b.confidence);
            double target = 0.0;
            if (auto_enabled_) {
                const bool voice_active = voice_db > -55.0 || level_db > -55.0;
                if (voice_active) {
                    const double baseline_target = rclamp(ratio_db * 0.45, 0.0, 24.0);
                    const double c = baseline_target < b.baseline_excess_db ? baseline_fall : baseline_rise;
                    b.baseline_excess_db = c * b.baseline_excess_db + (1.0 - c) * baseline_target;
                } else {
                    b.baseline_excess_db *= baseline_inactive;
                }
                const double cap_db = rmin(auto_cap, max_reduction_db_ * 0.75);
                const double conf_gain = confidence_gain(b.confidence, confidence_floor);
                const double over_db = rmax(ratio_db - b.baseline_excess_db - trigger_offset_db, 0.0);
                target = rclamp(over_db * slope * conf_gain, 0.0, cap_db);
            } else if (level_db > threshold_db_) {
                const double ratio_threshold_db = rclamp((threshold_db_ + 60.0) * 0.10, 0.0, 6.0);
                const double level_over = level_db - threshold_db_;
                const double ratio_over = ratio_db - ratio_threshold_db;
                if (ratio_over > 0.0) {
                    const double over_db = rmin(level_over, ratio_over);
                    const double conf_gain = confidence_gain(b.confidence, 0.22);
                    target = rclamp((1.0 - (1.0 / ratio_)) * over_db * conf_gain, 0.0, max_reduction_db_ * 0.75);
                }
            }
            targets[i] = target;
            target_sum += target;
        }
        if (target_sum > max_reduction_db_ && target_sum > 0.0) {
            const double scale = max_reduction_db_ / target_sum;
            for (double& t : targets) t *= scale;
        }
        float processed = input;
        double total_reduction = 0.0;
        for (size_t i = 0; i < 3; ++i) {
            Band& b = bands_[i];
            b.reduction_db = smooth(b.reduction_db, targets[i], attack_coeff_, release_coeff_);
            total_reduction += b.reduction_db;
            const double dyn_gain = -b.reduction_db;
            if (std::fabs(b.dynamic_eq.gain_db() - dyn_gain) > 0.001) b.dynamic_eq.set_gain_db_immediate(dyn_gain);
            processed = b.dynamic_eq.process_sample(processed);
        }
        current_reduction_db_ = rmin(total_reduction, max_reduction_db_);
        detector_confidence_ = rclamp(aggregate_conf, 0.0, 1.0);
        return processed;
    }
    void process_block_inplace(float* buf, size_t n) {  // deesser.rs:550-560
        if (!enabled_) {
            current_reduction_db_ = 0.0;
            detector_confidence_ = 0.0;
            return;
        }
        for (size_t i = 0; i < n; ++i) buf[i] = process_sample(buf[i]);
    }

  private:
    struct Band {  // deesser.rs:34-62
        double low_hz, high_hz;
        double env = 0, confidence = 0, baseline_excess_db = 0, reduction_db = 0;
        Biquad detector_hp, detector_lp, dynamic_eq;
        Band(double lo, double hi, double fs)
            : low_hz(lo), high_hz(hi),
              detector_hp(BiquadType::HighPass, lo, 0.0, 0.707, fs),
              detector_lp(BiquadType::LowPass, hi, 0.0, 0.707, fs),
              dynamic_eq(BiquadType::Peaking, center_hz(lo, hi), 0.0, dynamic_q(lo, hi), fs) {}
        void set_bounds(double lo, double hi) {  // deesser.rs:64-73
            low_hz = lo;
            high_hz = hi;
            detector_hp.set_frequency(lo);
            detector_lp.set_frequency(hi);
            dynamic_eq.set_frequency(center_hz(lo, hi));
            dynamic_eq.set_q(dynamic_q(lo, hi));
        }
    };
    static double smooth(double prev, double input, double attack, double release) {  // deesser.rs:147-154
        const double c = input > prev ? attack : release;
        return c * prev + (1.0 - c) * input;
    }
    static double lerp(double a, double b, double t) { return a + (b - a) * t; }
    static double norm_range(double v, double s, double e) { return rclamp((v - s) / (e - s), 0.0, 1.0); }
    static double confidence_gain(double conf, double floor) { return norm_range(conf, rclamp(floor, 0.0, 0.95), 1.0); }
    static double confidence_target(double level_db, double voice_db, double narrowness) {  // deesser.rs:171-219
        const double ratio_db = rmax(level_db - voice_db, 0.0);
        const double ratio_conf = norm_range(ratio_db, 1.5, 10.0);
        const double level_conf = norm_range(level_db, -62.0, -24.0);
        const double voice_conf = norm_range(voice_db, -58.0, -34.0);
        const double narrow_support = (ratio_db > 6.0 && level_db > -45.0) ? 0.75 : 0.0;
        const double voice_support = rmax(voice_conf, narrow_support);
        const double balance_conf = ratio_conf > 0.12 ? rmax(ratio_conf, voice_support * 0.65) : ratio_conf;
        const double broadband_penalty = lerp(0.35, 1.0, balance_conf);
        const double narrowness_gain = lerp(0.35, 1.0, norm_range(narrowness, 0.34, 0.68));
        return (0.62 * ratio_conf + 0.18 * level_conf + 0.20 * voice_support) * broadband_penalty * narrowness_gain;
    }
    void rebuild() {  // deesser.rs:226-240
        const double span = rmax(high_cut_hz_ - low_cut_hz_, 600.0);
        const double a = low_cut_hz_ + span / 3.0;
        const double b = low_cut_hz_ + span * 2.0 / 3.0;
        bands_[0].set_bounds(low_cut_hz_, a);
        bands_[1].set_bounds(a, b);
        bands_[2].set_bounds(b, high_cut_hz_);
    }

    bool enabled_ = false, auto_enabled_ = true;
    double auto_amount_ = 0.5, threshold_db_ = -28.0, ratio_ = 4.0;
    double attack_coeff_, release_coeff_, detector_attack_coeff_, detector_release_coeff_;
    double max_reduction_db_ = 6.0, current_reduction_db_ = 0.0, broadband_env_ = 0.0, detector_confidence_ = 0.0;
    double low_cut_hz_ = 4000.0, high_cut_hz_ = 11000.0;
    double sample_rate_;
    std::vector<Band> bands_;
};

// ---- dsp/loudness.rs over the `ebur128` crate 0.1.10 (momentary mode) ----------------------
// PARITY UNPINNED: the crate source is not in the reference tree.  Restated from ITU-R
// BS.1770-4 / libebur128's structure (which the crate ports): K-weighting = high shelf
// (f0 1681.974450955533 Hz, G 3.999843853973347 dB, Q 0.7071752369554196) cascaded with a
// high-pass (f0 38.13547087602444 Hz, Q 0.5003270373238773), coefficients re-derived for the
// sample rate and folded into one 4th-order direct-form-II section in f64; momentary loudness
// = -0.691 + 10 log10(mean square of the last 400 ms).
class LoudnessMeter {
  public:
    static bool supported_rate(uint32_t fs) {  // loudness.rs:37-42
        const uint32_t ok[] = {8000, 16000, 32000, 44100, 48000, 88200, 96000};
        return std::find(std::begin(ok), std::end(ok), fs) != std::end(ok);
    }
    explicit LoudnessMeter(uint32_t fs) : fs_(fs) {
        const double f0 = 1681.974450955533, G = 3.999843853973347, Q = 0.7071752369554196;
        double K = std::tan(kPi * f0 / fs);
        const double Vh = std::pow(10.0, G / 20.0);
        const double Vb = std::pow(Vh, 0.4996667741545416);
        double pb[3], pa[3] = {1.0, 0.0, 0.0};
        const double a0 = 1.0 + K / Q + K * K;
        pb[0] = (Vh + Vb * K / Q + K * K) / a0;
        pb[1] = 2.0 * (K * K - Vh) / a0;
        pb[2] = (Vh - Vb * K / Q + K * K) / a0;
        pa[1] = 2.0 * (K * K - 1.0) / a0;
        pa[2] = (1.0 - K / Q + K * K) / a0;
        const double f1 = 38.13547087602444, Q1 = 0.5003270373238773;
        K = std::tan(kPi * f1 / fs);
        const double rb[3] = {1.0, -2.0, 1.0};
        double ra[3] = {1.0, 0.0, 0.0};
        ra[1] = 2.0 * (K * K - 1.0) / (1.0 + K / Q1 + K * K);
        ra[2] = (1.0 - K / Q1 + K * K) / (1.0 + K / Q1 + K * K);
        b_[0] = pb[0] * rb[0];
        b_[1] = pb[0] * rb[1] + pb[1] * rb[0];
        b_[2] = pb[0] * rb[2] + pb[1] * rb[1] + pb[2] * rb[0];
        b_[3] = pb[1] * rb[2] + pb[2] * rb[1];
        b_[4] = pb[2] * rb[2];
        a_[0] = pa[0] * ra[0];
        a_[1] = pa[0] * ra[1] + pa[1] * ra[0];
        a_[2] = pa[0] * ra[2] + pa[1] * ra[1] + pa[2] * ra[0];
        a_[3] = pa[1] * ra[2] + pa[2] * ra[1];
        a_[4] = pa[2] * ra[2];
        window_ = static_cast<size_t>(fs) * 400 / 1000;
        ring_.assign(window_, 0.0);
    }
    void process(const float* samples, size_t n) {  // loudness.rs:113-131 + add_frames_f32
        for (size_t i = 0; i < n; ++i) {
            const double x = static_cast<double>(samples[i]);
            v_[0] = x - a_[1] * v_[1] - a_[2] * v_[2] - a_[3] * v_[3] - a_[4] * v_[4];
            const double y = b_[0] * v_[0] + b_[1] * v_[1] + b_[2] * v_[2] + b_[3] * v_[3] + b_[4] * v_[4];
            v_[4] = v_[3]; v_[3] = v_[2]; v_[2] = v_[1]; v_[1] = v_[0];
            for (double& v : v_) if (std::fabs(v) < std::numeric_limits<double>::min()) v = 0.0;
            ring_[pos_] = y;
            pos_ = (pos_ + 1) % window_;
            if (filled_ < window_) ++filled_;
        }
        // Momentary loudness over the last 400 ms (zeros before the ring has filled).
        double sum = 0.0;
        for (size_t i = 0; i < window_; ++i) sum += ring_[i] * ring_[i];
        const double energy = sum / static_cast<double>(window_);
        const double lufs = energy <= 0.0 ? -std::numeric_limits<double>::infinity()
                                          : 10.0 * std::log10(energy) - 0.691;
        current_lufs_ = static_cast<float>(lufs);
    }
    float loudness_momentary() const { return current_lufs_; }
    void reset() { *this = LoudnessMeter(fs_); }  // loudness.rs:146-154: a fresh meter, -100 LUFS

  private:
    uint32_t fs_;
    double b_[5], a_[5], v_[5] = {0, 0, 0, 0, 0};
    std::vector<double> ring_;
    size_t window_ = 0, pos_ = 0, filled_ = 0;
    float current_lufs_ = -100.0f;  // loudness.rs:105
};

// ---- dsp/compressor.rs ---------------------------------------------------------------------
struct AutoMakeupActivityInput {  // compressor.rs:31-36
    double vad_probability, vad_reliability, noise_floor_db, live_noise_reliability;
};

class Compressor {
  public:
    Compressor(double threshold_db, double ratio, double attack_ms, double release_ms, double makeup_gain_db,
               double knee_db, double sample_rate)  // compressor.rs:133-202
        : threshold_db_(threshold_db), ratio_(rmax(ratio, 1.0)), makeup_gain_db_(makeup_gain_db),
          knee_db_(rmax(knee_db, 0.0)), sample_rate_(sample_rate), base_release_ms_(release_ms),
          current_release_ms_(release_ms), target_release_ms_(release_ms), smoothed_makeup_gain_(makeup_gain_db) {
        attack_coeff_ = time_constant_to_coeff(attack_ms, sample_rate);
        release_coeff_ = time_constant_to_coeff(release_ms, sample_rate);
        detector_release_coeff_ = release_coeff_;
        rms_coeff_ = time_constant_to_coeff(20.0, sample_rate);
        release_smoothing_coeff_ = time_constant_to_coeff(100.0, sample_rate);
        makeup_smoothing_coeff_ = time_constant_to_coeff(200.0, sample_rate);
        speech_activity_smoothing_coeff_ = time_constant_to_coeff(200.0, sample_rate);
        makeup_silence_relax_coeff_ = time_constant_to_coeff(1500.0, sample_rate);
        makeup_gain_linear_ = db_to_linear(makeup_gain_db);
        const uint32_t fs_u32 = static_cast<uint32_t>(as_usize(sample_rate));  // `sample_rate as u32`
        if (LoudnessMeter::supported_rate(fs_u32)) meter_ = std::make_unique<LoudnessMeter>(fs_u32);
        sidechain_hp_coeff_ = sidechain_highpass_coeff(120.0, sample_rate);
    }
    void set_threshold(double v) { threshold_db_ = v; reset_adaptive_release_state(); }  // :210-213
    void set_ratio(double v) { ratio_ = rmax(v, 1.0); }
    void set_attack_time(double ms) { attack_coeff_ = time_constant_to_coeff(ms, sample_rate_); }
    void set_release_time(double ms) {  // :236-244
        base_release_ms_ = ms;
        if (!adaptive_release_) {
            current_release_ms_ = ms;
            target_release_ms_ = ms;
            release_coeff_ = time_constant_to_coeff(ms, sample_rate_);
        }
        detector_release_coeff_ = time_constant_to_coeff(ms, sample_rate_);
    }
    void set_adaptive_release(bool e) {  // :247-260
        adaptive_release_ = e;
        if (!e) {
            current_release_ms_ = base_release_ms_;
            target_release_ms_ = base_release_ms_;
            fast_release_env_db_ = current_gain_reduction_db_;
            slow_release_env_db_ = 0.0;
            release_coeff_ = time_constant_to_coeff(current_release_ms_, sample_rate_);
        } else {
            fast_release_env_db_ = current_gain_reduction_db_;
            slow_release_env_db_ = 0.0;
        }
    }
    void set_base_release_time(double ms) {  // :268-275
        base_release_ms_ = ms;
        if (!adaptive_release_) {
            current_release_ms_ = ms;
            target_release_ms_ = ms;
            release_coeff_ = time_constant_to_coeff(ms, sample_rate_);
        }
    }
    void reset_adaptive_release_state() { fast_release_env_db_ = current_gain_reduction_db_; slow_release_env_db_ = 0.0; }
    void set_makeup_gain(double db) {  // :294-300
        makeup_gain_db_ = db;
        makeup_gain_linear_ = db_to_linear(db);
        if (!auto_makeup_enabled_) smoothed_makeup_gain_ = db;
    }
    void set_enabled(bool e) { enabled_ = e; }
    bool is_enabled() const { return enabled_; }
    double current_gain_reduction() const { return current_gain_reduction_db_; }
    void set_auto_makeup_enabled(bool e) {  // :318-323
        auto_makeup_enabled_ = e && meter_ != nullptr;
        if (!e) smoothed_makeup_gain_ = makeup_gain_db_;
    }
    void set_target_lufs(double t) { target_lufs_ = rclamp(t, -24.0, -12.0); }
    double current_makeup_gain() const { return smoothed_makeup_gain_; }
    void set_noise_reference_reliability(double r) { noise_reference_reliability_ = std::isfinite(r) ? rclamp(r, 0.0, 1.0) : 0.0; }
    double auto_makeup_activity() const { return speech_activity_score_; }
    double auto_makeup_activity_reliability() const { return auto_makeup_activity_reliability_; }
    void set_sidechain_highpass_enabled(bool e) {  // :366-371
        if (sidechain_hp_enabled_ != e) reset_sidechain_state();
        sidechain_hp_enabled_ = e;
    }
    double plosive_ratio() const { return plosive_ratio_; }
    double current_release_time() const { return current_release_ms_; }

    static double sidechain_highpass_coeff(double cutoff_hz, double sample_rate) {  // :390-394
        const double c = rclamp(cutoff_hz, 20.0, sample_rate * 0.45);
        const double omega = 2.0 * kPi * c / rmax(sample_rate, 1.0);
        return 1.0 / (1.0 + omega);
    }
    // :657-678
    double compute_gain_reduction(double detector_db) const {
        const double comp_factor = 1.0 - 1.0 / ratio_;
        if (knee_db_ <= 0.0) {
            if (detector_db <= threshold_db_) return 0.0;
            return (detector_db - threshold_db_) * comp_factor;
        }
        const double knee_half = knee_db_ / 2.0;
        const double knee_start = threshold_db_ - knee_half;
        const double knee_end = threshold_db_ + knee_half;
        if (detector_db <= knee_start) return 0.0;
        if (detector_db >= knee_end) return (detector_db - threshold_db_) * comp_factor;
        const double x = detector_db - knee_start;
        return comp_factor * x * x / (2.0 * knee_db_);
    }
    static double blended_detector_db(double peak_db, double rms_db) {  // :681-686
        const double blended = 0.6 * db_to_linear(peak_db) + 0.4 * db_to_linear(rms_db);
        return linear_to_db(blended, 1e-10);
    }
    void set_limiter_feedback_gain_reduction_db(double db) { limiter_feedback_gr_db_ = rclamp(db, 0.0, 24.0); }  // :385-387
    // the reference's unit tests call these private methods directly (compressor.rs:1074-1095,1169-1245)
    void test_update_auto_makeup_gain(double activity, double reliability, size_t elapsed) {
        update_auto_makeup_gain(activity, reliability, elapsed);
    }
    void test_estimate_activity(double rms_db, const AutoMakeupActivityInput* ev, double* activity, double* reliability) const {
        const Activity a = estimate_activity(rms_db, ev);
        *activity = a.activity;
        *reliability = a.reliability;
    }
    static double test_speech_activity_from_rms_db(double rms_db) { return speech_activity_from_rms_db(rms_db); }
    float process_sample(float input) { return process_sample_impl(input, true); }
    void process_block_inplace(float* buf, size_t n) { process_block_with_activity(buf, n, nullptr); }
    void process_block_with_activity(float* buf, size_t n, const AutoMakeupActivityInput* evidence) {  // :700-722
        if (!enabled_) {
            current_gain_reduction_db_ = 0.0;
            return;
        }
        const Activity act = estimate_activity(block_rms_db(buf, n), evidence);
        for (size_t i = 0; i < n; ++i) buf[i] = process_sample_impl(buf[i], false);
        if (act.activity > 0.20 && act.reliability >= 0.35) {
            if (meter_) meter_->process(buf, n);
        }
        update_auto_makeup_gain(act.activity, act.reliability, n);
    }

  private:
    struct Activity { double activity, reliability; };
    static double speech_activity_from_rms_db(double rms_db) {  // :507-514
        if (!(rms_db >= -55.0 && rms_db <= -6.0)) return 0.0;
        const double onset = rclamp((rms_db - -55.0) / 12.0, 0.0, 1.0);
        const double overload = rclamp((-6.0 - rms_db) / 6.0, 0.0, 1.0);
        return rmin(onset, overload);
    }
    static bool finite_unit(double v, double& out) {
        if (!std::isfinite(v)) return false;
        out = rclamp(v, 0.0, 1.0);
        return true;
    }
    static double smoothstep(double e0, double e1, double v) {  // :520-526
        if (!std::isfinite(v) || !std::isfinite(e0) || !std::isfinite(e1) || e1 <= e0) return 0.0;
        const double t = rclamp((v - e0) / (e1 - e0), 0.0, 1.0);
        return t * t * (3.0 - 2.0 * t);
    }
    Activity estimate_activity(double rms_db, const AutoMakeupActivityInput* ev) const {  // :528-581
        const double absolute = speech_activity_from_rms_db(rms_db);
        if (!ev) return {absolute, 1.0};
        double vad_rel = 0.0;
        if (!finite_unit(ev->vad_reliability, vad_rel)) vad_rel = 0.0;
        double vad_p = 0.0;
        if (!finite_unit(ev->vad_probability, vad_p)) { vad_rel = 0.0; vad_p = 0.0; }
        double cfg_rel = 0.0;
        if (!finite_unit(noise_reference_reliability_, cfg_rel)) cfg_rel = 0.0;
        double live_rel = 0.0;
        if (!finite_unit(ev->live_noise_reliability, live_rel)) live_rel = 0.0;
        double noise_rel = cfg_rel > 0.0 ? rmin(live_rel, cfg_rel) : live_rel;
        double relative = 0.0;
        if (std::isfinite(ev->noise_floor_db) && ev->noise_floor_db >= -120.0 && ev->noise_floor_db <= 0.0) {
            relative = smoothstep(ev->noise_floor_db + 3.0, ev->noise_floor_db + 15.0, rms_db);
        } else {
            noise_rel = 0.0;
        }
        const double fallback = noise_rel * relative + (1.0 - noise_rel) * absolute;
        const double activity = vad_rel * vad_p + (1.0 - vad_rel) * fallback;
        const double reliability = rmax(vad_rel, 0.75 * noise_rel);
        return {rclamp(activity, 0.0, 1.0), rclamp(reliability, 0.0, 1.0)};
    }
    static double block_rms_db(const float* buf, size_t n) {  // :583-596
        if (n == 0) return -120.0;
        double sum = 0.0;
        for (size_t i = 0; i < n; ++i) {
            const double s = static_cast<double>(buf[i]);
            sum += s * s;
        }
        return linear_to_db(std::sqrt(sum / static_cast<double>(n)), 1e-10);
    }
    void update_auto_makeup_gain(double speech_activity, double reliability, size_t elapsed) {  // :598-653
        const double n = static_cast<double>(std::max<size_t>(elapsed, 1));
        const double makeup_coeff = std::pow(makeup_smoothing_coeff_, n);
        const double relax_coeff = std::pow(makeup_silence_relax_coeff_, n);
        if (!auto_makeup_enabled_) {
            const double target = makeup_gain_db_;
            const double diff = target - smoothed_makeup_gain_;
            if (std::fabs(diff) > 0.1)
                smoothed_makeup_gain_ = makeup_coeff * smoothed_makeup_gain_ + (1.0 - makeup_coeff) * target;
            else
                smoothed_makeup_gain_ = target;
            return;
        }
        if (!meter_) return;
        current_lufs_ = static_cast<double>(meter_->loudness_momentary());
        const double activity_coeff = std::pow(speech_activity_smoothing_coeff_, n);
        speech_activity_score_ = activity_coeff * speech_activity_score_ + (1.0 - activity_coeff) * rclamp(speech_activity, 0.0, 1.0);
        auto_makeup_activity_reliability_ = rclamp(reliability, 0.0, 1.0);
        if (speech_activity_score_ < 0.20) {
            smoothed_makeup_gain_ = relax_coeff * smoothed_makeup_gain_ + (1.0 - relax_coeff) * makeup_gain_db_;
            return;
        }
        if (auto_makeup_activity_reliability_ < 0.35) {
            const double cap = makeup_gain_db_ + 3.0 * (auto_makeup_activity_reliability_ / 0.35);
            if (smoothed_makeup_gain_ > cap)
                smoothed_makeup_gain_ = makeup_coeff * smoothed_makeup_gain_ + (1.0 - makeup_coeff) * cap;
            return;
        }
        const double required = target_lufs_ - current_lufs_;
        const double reliability_cap = rclamp(12.0 * auto_makeup_activity_reliability_, 3.0, 12.0);
        const double headroom_cap = rclamp(12.0 - limiter_feedback_gr_db_ * 2.0, 0.0, reliability_cap);
        const double clamped = rclamp(required, 0.0, headroom_cap);
        const double diff = clamped - smoothed_makeup_gain_;
        if (std::fabs(diff) > 0.1)
            smoothed_makeup_gain_ = makeup_coeff * smoothed_makeup_gain_ + (1.0 - makeup_coeff) * clamped;
        else
            smoothed_makeup_gain_ = clamped;
    }
    void reset_sidechain_state() {  // :397-404
        sc_prev_in_ = 0.0; sc_prev_out_ = 0.0;
        low_env_sq_ = 0.0; voiced_env_sq_ = 0.0; presence_env_sq_ = 0.0; plosive_ratio_ = 0.0;
    }
    double sidechain(double x) {  // :407-417
        if (!sidechain_hp_enabled_) return x;
        const double y = sidechain_hp_coeff_ * (sc_prev_out_ + x - sc_prev_in_);
        sc_prev_in_ = x;
        sc_prev_out_ = y;
        return y;
    }
    double band_metrics(double full, double det) {  // :420-450
        if (!sidechain_hp_enabled_) {
            plosive_ratio_ = 0.0;
            return 1.0;
        }
        const double low = full - det;
        const double voiced = det;
        const double presence = 0.65 * det + 0.35 * (det - low);
        const double c = time_constant_to_coeff(18.0, sample_rate_);
        low_env_sq_ = c * low_env_sq_ + (1.0 - c) * low * low;
        voiced_env_sq_ = c * voiced_env_sq_ + (1.0 - c) * voiced * voiced;
        presence_env_sq_ = c * presence_env_sq_ + (1.0 - c) * presence * presence;
        const double low_rms = std::sqrt(low_env_sq_);
        const double voiced_rms = rmax(std::sqrt(voiced_env_sq_), 1e-8);
        const double presence_rms = std::sqrt(presence_env_sq_);
        plosive_ratio_ = rclamp(low_rms / voiced_rms, 0.0, 32.0);
        const double amount = rclamp((plosive_ratio_ - 1.25) / (5.0 - 1.25), 0.0, 1.0);
        const double penalty = 1.0 - amount * (1.0 - 0.35);
        const double presence_ratio = rclamp(presence_rms / voiced_rms, 0.0, 4.0);
        const double weight = 1.0 + 0.18 * rclamp(presence_ratio - 0.75, 0.0, 1.0);
        return rclamp(penalty * weight, 0.35, 1.15);
    }
    void update_release_meter() {  // :452-466
        if (!adaptive_release_) {
            target_release_ms_ = base_release_ms_;
            return;
        }
        const double sustained = rclamp(slow_release_env_db_ / (3.0 + 3.0), 0.0, 1.0);
        const double transient = rclamp((fast_release_env_db_ - slow_release_env_db_) / (3.0 + 4.0), 0.0, 1.0);
        const double syllabic = rclamp(sustained * sustained * (1.0 - 0.35 * transient), 0.0, 1.0);
        target_release_ms_ = 50.0 + syllabic * (400.0 - 50.0);
    }
    void smooth_gain_reduction(double target) {  // :468-505
        if (!adaptive_release_) {
            const double c = target > current_gain_reduction_db_ ? attack_coeff_ : release_coeff_;
            current_gain_reduction_db_ = c * current_gain_reduction_db_ + (1.0 - c) * target;
            fast_release_env_db_ = current_gain_reduction_db_;
            slow_release_env_db_ = 0.0;
            return;
        }
        const double fast_c = time_constant_to_coeff(50.0, sample_rate_);
        const double charge_c = time_constant_to_coeff(250.0, sample_rate_);
        const double slow_c = time_constant_to_coeff(400.0, sample_rate_);
        if (target > current_gain_reduction_db_)
            fast_release_env_db_ = attack_coeff_ * current_gain_reduction_db_ + (1.0 - attack_coeff_) * target;
        else
            fast_release_env_db_ = fast_c * fast_release_env_db_ + (1.0 - fast_c) * target;
        if (target > 3.0)
            slow_release_env_db_ = charge_c * slow_release_env_db_ + (1.0 - charge_c) * target;
        else
            slow_release_env_db_ *= slow_c;
        current_gain_reduction_db_ = rmax(fast_release_env_db_, slow_release_env_db_);
    }
    float process_sample_impl(float input, bool update_makeup) {  // :725-774
        if (!enabled_) {
            current_gain_reduction_db_ = 0.0;
            return input;
        }
        const double x = static_cast<double>(input);
        const double det = sidechain(x);
        const double weight = band_metrics(x, det);
        const double inst_peak_db = linear_to_db(std::fabs(det), 1e-10);
        const double pc = inst_peak_db > peak_envelope_db_ ? attack_coeff_ : detector_release_coeff_;
        peak_envelope_db_ = pc * peak_envelope_db_ + (1.0 - pc) * inst_peak_db;
        const double sq = det * det;
        rms_envelope_sq_ = rms_coeff_ * rms_envelope_sq_ + (1.0 - rms_coeff_) * sq;
        const double rms_db = linear_to_db(std::sqrt(rms_envelope_sq_), 1e-10);
        const double detector_db = blended_detector_db(peak_envelope_db_, rms_db) + linear_to_db(weight, 1e-10);
        update_release_meter();
        const double diff = target_release_ms_ - current_release_ms_;
        if (std::fabs(diff) > 1.0)
            current_release_ms_ = release_smoothing_coeff_ * current_release_ms_ + (1.0 - release_smoothing_coeff_) * target_release_ms_;
        else
            current_release_ms_ = target_release_ms_;
        release_coeff_ = time_constant_to_coeff(current_release_ms_, sample_rate_);
        const double target_gr = compute_gain_reduction(detector_db);
        smooth_gain_reduction(target_gr);
        if (update_makeup) update_auto_makeup_gain(speech_activity_from_rms_db(detector_db), 1.0, 1);
        const double gain = db_to_linear(-current_gain_reduction_db_) * db_to_linear(smoothed_makeup_gain_);
        return static_cast<float>(x * gain);
    }

    double threshold_db_, ratio_, attack_coeff_, release_coeff_, detector_release_coeff_;
    double makeup_gain_db_, makeup_gain_linear_, knee_db_;
    double peak_envelope_db_ = -120.0, rms_envelope_sq_ = 0.0, rms_coeff_;
    double current_gain_reduction_db_ = 0.0, sample_rate_;
    bool enabled_ = true, adaptive_release_ = false;
    double base_release_ms_, current_release_ms_, target_release_ms_, release_smoothing_coeff_;
    double fast_release_env_db_ = 0.0, slow_release_env_db_ = 0.0;
    std::unique_ptr<LoudnessMeter> meter_;
    bool auto_makeup_enabled_ = false;
    double target_lufs_ = -18.0, smoothed_makeup_gain_, makeup_smoothing_coeff_, current_lufs_ = -100.0;
    double speech_activity_score_ = 0.0, speech_activity_smoothing_coeff_;
    double auto_makeup_activity_reliability_ = 0.0, noise_reference_reliability_ = 0.0, makeup_silence_relax_coeff_;
    bool sidechain_hp_enabled_ = false;
    double sidechain_hp_coeff_, sc_prev_in_ = 0.0, sc_prev_out_ = 0.0;
    double low_env_sq_ = 0.0, voiced_env_sq_ = 0.0, presence_env_sq_ = 0.0, plosive_ratio_ = 0.0;
    double limiter_feedback_gr_db_ = 0.0;
};

// ---- dsp/limiter.rs ----------------------------------------------------------------------------
class Limiter {
  public:
    Limiter(double ceiling_db, double release_ms, double sample_rate, double lookahead_ms = 2.0)  // :105-131
        : ceiling_db_(ceiling_db), sample_rate_(sample_rate) {
        ceiling_linear_ = db_to_linear(ceiling_db);
        release_coeff_ = time_constant_to_coeff(release_ms, sample_rate);
        lookahead_ = lookahead_samples_for(lookahead_ms, sample_rate);
        delay_.assign(lookahead_, 0.0f);
    }
    static size_t lookahead_samples_for(double ms, double fs) {  // :113-115
        return std::clamp<size_t>(as_usize(std::round(rclamp(ms, 0.1, 10.0) / 1000.0 * fs)), 1, 1024);
    }
    void set_ceiling(double db) { ceiling_db_ = rmin(db, 0.0); ceiling_linear_ = db_to_linear(ceiling_db_); }  // :139-142
    double ceiling_db() const { return ceiling_db_; }
    void set_release_time(double ms) { release_coeff_ = time_constant_to_coeff(ms, sample_rate_); }
    void set_lookahead_ms(double ms) {  // :157-166
        const size_t n = lookahead_samples_for(ms, sample_rate_);
        if (n != lookahead_) {
            lookahead_ = n;
            delay_.resize(n, 0.0f);
            reset();
        }
    }
    size_t lookahead_samples() const { return lookahead_; }
    void set_enabled(bool e) {  // :179-184
        if (enabled_ != e) reset();
        enabled_ = e;
    }
    bool is_enabled() const { return enabled_; }
    double peak_gain_reduction_and_reset() { const double p = peak_gr_db_; peak_gr_db_ = 0.0; return p; }
    float process_sample(float input) {  // :246-284
        if (!enabled_) return input;
        const double delayed = static_cast<double>(delay_[write_idx_]);
        const double in_abs = std::fabs(static_cast<double>(input));
        const double window_peak = queue_.empty() ? 0.0 : queue_.front().second;
        const double peak = rmax(window_peak, in_abs);
        delay_[write_idx_] = input;
        push(in_abs);
        write_idx_ = (write_idx_ + 1) % lookahead_;
        const double target = peak > ceiling_linear_ ? ceiling_linear_ / peak : 1.0;
        if (target < gain_)
            gain_ = target;
        else
            gain_ = release_coeff_ * gain_ + (1.0 - release_coeff_) * target;
        const double reduction_db = gain_ < 1.0 ? -linear_to_db(gain_, 1e-10) : 0.0;
        if (reduction_db > peak_gr_db_) peak_gr_db_ = reduction_db;
        const double limited = delayed * gain_;
        return static_cast<float>(rclamp(limited, -ceiling_linear_, ceiling_linear_));
    }
    void process_block_inplace(float* buf, size_t n) {
        if (!enabled_) return;
        for (size_t i = 0; i < n; ++i) buf[i] = process_sample(buf[i]);
    }
    void reset() {  // :297-304
        gain_ = 1.0;
        peak_gr_db_ = 0.0;
        next_index_ = 0;
        write_idx_ = 0;
        std::fill(delay_.begin(), delay_.end(), 0.0f);
        queue_.clear();
    }

  private:
    void push(double v) {  // :216-237 monotonic deque over the last `lookahead_` samples
        while (!queue_.empty() && !(queue_.back().second > v)) queue_.pop_back();
        if (queue_.size() == 1024) queue_.pop_front();
        queue_.emplace_back(next_index_, v);
        next_index_ += 1;
        const uint64_t oldest = next_index_ >= lookahead_ ? next_index_ - lookahead_ : 0;
        while (!queue_.empty() && queue_.front().first < oldest) queue_.pop_front();
    }
    double ceiling_db_, ceiling_linear_, release_coeff_, gain_ = 1.0, peak_gr_db_ = 0.0, sample_rate_;
    size_t lookahead_;
    std::vector<float> delay_;
    std::deque<std::pair<uint64_t, double>> queue_;
    uint64_t next_index_ = 0;
    size_t write_idx_ = 0;
    bool enabled_ = true;
};

// ---- dsp/true_peak.rs -------------------------------------------------------------------------
inline const float (*true_peak_fir())[32] {
    static const float table[4][32] = {
#include "true_peak_fir.inc"
    };
    return table;
}

class Bandlimited4xPeak {  // true_peak.rs:156-187
  public:
    float observe(float sample) {
        std::memmove(history_ + 1, history_, 31 * sizeof(float));
        history_[0] = sample;
        const float (*fir)[32] = true_peak_fir();
        float peak = std::fabs(sample);
        for (int p = 0; p < 4; ++p) {
            float acc = 0.0f;
            for (int k = 0; k < 32; ++k) acc = std::fmaf(fir[p][k], history_[k], acc);
            peak = rmaxf(peak, std::fabs(acc));
        }
        return peak;
    }
    void reset() { std::memset(history_, 0, sizeof history_); }

  private:
    float history_[32] = {};
};

class TruePeakDetector {  // true_peak.rs:189-232
  public:
    float process_block(const float* samples, size_t n) {
        float peak = 0.0f;
        for (size_t i = 0; i < n; ++i) {
            const float s = std::isfinite(samples[i]) ? samples[i] : 0.0f;
            peak = rmaxf(peak, os_.observe(s));
        }
        return peak;
    }

  private:
    Bandlimited4xPeak os_;
};

struct TruePeakLimiterBlockStats {
    uint64_t limited_events = 0;
    float input_true_peak = 0, output_true_peak = 0, max_gain_reduction_db = 0;
};

class TruePeakLimiter {  // true_peak.rs:252-392
  public:
    TruePeakLimiter(float sample_rate, float ceiling_db, float release_ms) : sample_rate_(rmaxf(sample_rate, 1.0f)) {
        ceiling_linear_ = static_cast<float>(db_to_linear(static_cast<double>(ceiling_db)));
        // The constructor derives the coefficient from the UNCLAMPED sample rate first, then
        // set_release_ms overwrites it (true_peak.rs:267-283); only the latter survives.
        set_release_ms(release_ms);
    }
    void set_ceiling_linear(float c) { ceiling_linear_ = rclampf(c, 0.000001f, 1.0f); }
    void set_release_ms(float ms) {
        release_coeff_ = static_cast<float>(
            time_constant_to_coeff(static_cast<double>(rclampf(ms, 5.0f, 500.0f)), static_cast<double>(sample_rate_)));
    }
    float current_gain_reduction_db() const {
        return gain_ >= 1.0f ? 0.0f : -20.0f * std::log10(rmaxf(gain_, 1e-10f));
    }
    TruePeakLimiterBlockStats process_block_inplace(float* samples, size_t n) {  // :337-378
        TruePeakLimiterBlockStats stats;
        bool limited = false;
        for (size_t i = 0; i < n; ++i) {
            const float input = std::isfinite(samples[i]) ? samples[i] : 0.0f;
            const float delayed = delay_[write_idx_];
            delay_[write_idx_] = input;
            write_idx_ = (write_idx_ + 1) % 20;
            const float itp = in_os_.observe(input);
            stats.input_true_peak = rmaxf(stats.input_true_peak, itp);
            const float target = itp > ceiling_linear_ ? rclampf((ceiling_linear_ * 0.999f) / itp, 0.0f, 1.0f) : 1.0f;
            if (target < gain_) {
                gain_ = target;
                limited = true;
            } else {
                gain_ = release_coeff_ * gain_ + (1.0f - release_coeff_) * target;
            }
            const float red = current_gain_reduction_db();
            peak_gr_db_ = rmaxf(peak_gr_db_, red);
            stats.max_gain_reduction_db = rmaxf(stats.max_gain_reduction_db, red);
            float out = rclampf(delayed * gain_, -ceiling_linear_, ceiling_linear_);
            if (!std::isfinite(out)) out = 0.0f;
            stats.output_true_peak = rmaxf(stats.output_true_peak, out_os_.observe(out));
            samples[i] = out;
        }
        stats.limited_events = limited ? 1 : 0;
        return stats;
    }

  private:
    float ceiling_linear_, release_coeff_, gain_ = 1.0f;
    float delay_[20] = {};
    size_t write_idx_ = 0;
    Bandlimited4xPeak in_os_, out_os_;
    float peak_gr_db_ = 0.0f, sample_rate_;
};

// ---- audio/processor/routing.rs: input stage ------------------------------------------------------
struct InputPreFilterState { float dc_x1 = 0.0f, dc_y1 = 0.0f; };  // routing.rs:8-12

// routing.rs:826-843, consts processor.rs:74-76
inline void apply_input_pre_filter(float* buf, size_t n, InputPreFilterState& st, Biquad& hp, bool fixed_highpass) {
    for (size_t i = 0; i < n; ++i) {
        const float in = buf[i];
        const float out = in - st.dc_x1 + 0.995f * st.dc_y1;
        st.dc_x1 = in;
        st.dc_y1 = out;
        buf[i] = fixed_highpass ? hp.process_sample(out) : out;
    }
}

enum class CleanupMode : uint8_t { Off = 0, Gentle = 1, Strong = 2 };

inline float wrap_phase(float p) {  // routing.rs:598-607
    const float two_pi = 2.0f * kPiF;
    while (p > kPiF) p -= two_pi;
    while (p < -kPiF) p += two_pi;
    return p;
}
inline float smooth_toward(float cur, float target, float attack, float release) {  // routing.rs:642-646
    const float c = target > cur ? attack : release;
    return cur + c * (target - cur);
}

struct HumBin {  // routing.rs:55-110
    float cos_phase = 1.0f, sin_phase = 0.0f, cos_step, sin_step, i_acc = 0.0f, q_acc = 0.0f;
    HumBin(float frequency_hz, float sample_rate) {
        const float omega = 2.0f * kPiF * frequency_hz / rmaxf(sample_rate, 1.0f);
        cos_step = std::cos(omega);
        sin_step = std::sin(omega);
    }
    void analyze(float s) {
        i_acc += s * cos_phase;
        q_acc += s * sin_phase;
        const float nc = cos_phase * cos_step - sin_phase * sin_step;
        const float ns = sin_phase * cos_step + cos_phase * sin_step;
        cos_phase = nc;
        sin_phase = ns;
    }
    std::pair<float, float> power_phase_and_reset(size_t window) {
        const float n = static_cast<float>(std::max<size_t>(window, 1));
        const float power = (i_acc * i_acc + q_acc * q_acc) * (2.0f / (n * n));
        const float phase = std::atan2(q_acc, i_acc);
        i_acc = 0.0f;
        q_acc = 0.0f;
        const float norm = std::sqrt(cos_phase * cos_phase + sin_phase * sin_phase);
        if (norm > 1.0e-6f) {
            cos_phase /= norm;
            sin_phase /= norm;
        }
        return {power, phase};
    }
};

struct NotchFilter {  // routing.rs:117-157
    float b0, b1, b2, a1, a2, z1 = 0.0f, z2 = 0.0f;
    NotchFilter(float f, float q, float fs) {
        const float omega = 2.0f * kPiF * f / rmaxf(fs, 1.0f);
        const float sn = std::sin(omega), cs = std::cos(omega);
        const float alpha = sn / (2.0f * rmaxf(q, 1.0f));
        const float a0 = 1.0f + alpha;
        b0 = 1.0f / a0;
        b1 = -2.0f * cs / a0;
        b2 = 1.0f / a0;
        a1 = -2.0f * cs / a0;
        a2 = (1.0f - alpha) / a0;
    }
    float process(float x) {
        const float y = b0 * x + z1;
        z1 = b1 * x - a1 * y + z2;
        z2 = b2 * x - a2 * y;
        return y;
    }
};

struct SmoothNotch {  // routing.rs:160-217
    NotchFilter active, pending;
    float frequency_hz, pending_frequency_hz;
    size_t fade_total, fade_remaining = 0;
    float sample_rate, q;
    SmoothNotch(float f, float q_, float fs)
        : active(f, q_, fs), pending(active), frequency_hz(f), pending_frequency_hz(f), sample_rate(fs), q(q_) {
        fade_total = as_usize(static_cast<double>(rmaxf(std::round(fs * 0.020f), 1.0f)));
    }
    void retune(float f) {
        f = rclampf(f, 20.0f, sample_rate * 0.45f);
        if (std::fabs(f - pending_frequency_hz) < 0.15f) return;
        pending = NotchFilter(f, q, sample_rate);
        pending_frequency_hz = f;
        fade_remaining = fade_total;
    }
    float process(float x) {
        const float a = active.process(x);
        if (fade_remaining == 0) return a;
        const float p = pending.process(x);
        const float fade = static_cast<float>(fade_total - fade_remaining + 1) / static_cast<float>(fade_total);
        const float out = a + (p - a) * fade;
        fade_remaining -= 1;
        if (fade_remaining == 0) {
            active = pending;
            frequency_hz = pending_frequency_hz;
        }
        return out;
    }
};

class AdaptiveInputCleanup {  // routing.rs:219-596 (fresh state as in processor/tests.rs:500-549)
  public:
    explicit AdaptiveInputCleanup(float fs)
        : fs_(fs), highpass_(BiquadType::HighPass, 80.0, 0.0, 0.707, static_cast<double>(fs)),
          hum_notch_(55.0f, 36.0f, fs), harmonic_notch_(110.0f, 36.0f, fs) {
        for (int i = 0; i < 13; ++i) {
            bins_.emplace_back(49.0f + static_cast<float>(i) * 1.0f, fs);
            harm_bins_.emplace_back(2.0f * (49.0f + static_cast<float>(i) * 1.0f), fs);
        }
        window_samples_ = as_usize(static_cast<double>(rmaxf(std::round(fs * 0.25f), 1.0f)));
    }
    void set_mode(CleanupMode m) { mode_ = m; }  // fresh object: no dynamic-state reset needed
    bool enabled() const { return mode_ != CleanupMode::Off; }
    float hum_line_hz() const { return hum_line_hz_; }
    bool hum_detected() const { return hum_detected_; }
    bool rumble_detected() const { return rumble_detected_; }
    float selected_high_pass_hz() const { return selected_hp_hz_; }
    uint32_t hum_hold_samples() const { return hum_hold_; }
    bool hum_phase_valid() const { return phase_valid_; }

    void analyze_input(const float* buf, size_t n) {  // routing.rs:336-403
        if (!enabled()) return;
        const float lowpass_coeff = rclampf(2.0f * kPiF * 150.0f / fs_, 0.0f, 1.0f);
        const float fast_attack = 0.08f, fast_release = 0.006f, slow_coeff = 0.0012f, broadband_coeff = 0.02f;
        for (size_t i = 0; i < n; ++i) {
            const float s = buf[i];
            total_energy_ += s * s;
            for (auto& b : bins_) b.analyze(s);
            for (auto& b : harm_bins_) b.analyze(s);
            window_pos_ += 1;
            if (window_pos_ >= window_samples_) finish_window();
            lowpass_ += lowpass_coeff * (s - lowpass_);
            const float low_abs = std::fabs(lowpass_);
            const float lc = low_abs > low_env_ ? fast_attack : fast_release;
            low_env_ += lc * (low_abs - low_env_);
            slow_low_env_ += slow_coeff * (low_abs - slow_low_env_);
            broadband_env_ += broadband_coeff * (std::fabs(s) - broadband_env_);
            const float burst_ratio = low_env_ / rmaxf(slow_low_env_, 0.006f);
            const float low_dominance = low_env_ / rmaxf(broadband_env_, 0.01f);
            const float threshold = mode_ == CleanupMode::Gentle ? 0.055f : 0.035f;
            const float ratio_threshold = mode_ == CleanupMode::Gentle ? 2.8f : 2.1f;
            const bool startup_burst = windows_observed_ == 0 && low_env_ > 0.45f;
            const bool established = windows_observed_ > 0 && slow_low_env_ > 0.012f;
            if ((startup_burst || established) && hum_hold_ == 0 && candidate_windows_ == 0 && low_env_ > threshold &&
                burst_ratio > ratio_threshold && low_dominance > 0.62f) {
                rumble_hold_ = mode_ == CleanupMode::Gentle ? as_u32(std::round(fs_ * 0.18f)) : as_u32(std::round(fs_ * 0.30f));
            } else {
                rumble_hold_ = rumble_hold_ > 0 ? rumble_hold_ - 1 : 0;
            }
            hum_hold_ = hum_hold_ > 0 ? hum_hold_ - 1 : 0;
        }
    }
    void process_block(float* buf, size_t n) {  // routing.rs:534-596
        if (!enabled()) return;
        hum_detected_ = hum_hold_ > 0;
        rumble_detected_ = rumble_hold_ > 0;
        if (rumble_detected_)
            selected_hp_hz_ = mode_ == CleanupMode::Gentle ? 100.0f : 120.0f;
        else
            selected_hp_hz_ = 80.0f;
        if (std::fabs(selected_hp_hz_ - highpass_hz_) > 0.5f) {
            highpass_.set_frequency(static_cast<double>(selected_hp_hz_));
            highpass_hz_ = selected_hp_hz_;
        }
        const float attack = mode_ == CleanupMode::Gentle ? 0.22f : 0.34f;
        const float release = 0.035f;
        const float target_hum = hum_detected_ ? (mode_ == CleanupMode::Gentle ? 0.55f : 0.85f) : 0.0f;
        const float target_harm = hum_detected_ ? (mode_ == CleanupMode::Strong ? 0.60f : 0.0f) : 0.0f;
        hum_strength_ = smooth_toward(hum_strength_, target_hum, attack, release);
        harmonic_strength_ = smooth_toward(harmonic_strength_, target_harm, attack, release);
        if (hum_line_hz_ > 0.0f) {
            hum_notch_.retune(hum_line_hz_);
            harmonic_notch_.retune(hum_line_hz_ * 2.0f);
        }
        for (size_t i = 0; i < n; ++i) {
            float y = buf[i];
            const float pn = hum_notch_.process(y);
            y += (pn - y) * rclampf(hum_strength_, 0.0f, 1.0f);
            const float hn = harmonic_notch_.process(y);
            y += (hn - y) * rclampf(harmonic_strength_, 0.0f, 1.0f);
            y = highpass_.process_sample(y);
            buf[i] = y;
        }
    }

  private:
    void finish_window() {  // routing.rs:405-532
        float best_f = 0.0f, best_primary = 0.0f, best_harm = 0.0f, best_score = 0.0f, best_phase = 0.0f;
        float primary[13];
        for (int i = 0; i < 13; ++i) {
            auto [pp, ph] = bins_[i].power_phase_and_reset(window_samples_);
            const float hp = harm_bins_[i].power_phase_and_reset(window_samples_).first;
            primary[i] = pp;
            const float score = pp + hp * 0.65f;
            if (score > best_score) {
                best_score = score;
                best_primary = pp;
                best_harm = hp;
                best_f = 49.0f + static_cast<float>(i) * 1.0f;
                best_phase = ph;
            }
        }
        const float total_power = total_energy_ / static_cast<float>(std::max<size_t>(window_samples_, 1)) + 1.0e-9f;
        window_pos_ = 0;
        windows_observed_ = windows_observed_ == 0xffffffffu ? windows_observed_ : windows_observed_ + 1;
        total_energy_ = 0.0f;
        const float primary_ratio = best_primary / total_power;
        const float harmonic_ratio = best_harm / total_power;
        const float ratio_threshold = mode_ == CleanupMode::Gentle ? 0.075f : 0.040f;
        const float power_threshold = mode_ == CleanupMode::Gentle ? 1.8e-5f : 8.0e-6f;
        const bool candidate = (best_primary > power_threshold || best_harm > power_threshold * 0.70f) &&
                               (primary_ratio > ratio_threshold || harmonic_ratio > ratio_threshold * 0.85f) &&
                               best_f > 0.0f;
        if (candidate) {
            candidate_windows_ = static_cast<uint8_t>(std::min<int>(candidate_windows_ + 1, 3));
        } else {
            candidate_windows_ = 0;
            phase_valid_ = false;
        }
        if (candidate_windows_ >= 2) {
            hum_hold_ = as_u32(std::round(fs_ * 0.75f));
            const float idx_f = rclampf(std::round((best_f - 49.0f) / 1.0f), 0.0f, 12.0f);
            const size_t idx = static_cast<size_t>(idx_f);
            float offset = 0.0f;
            if (idx > 0 && idx + 1 < 13) {
                const float left = std::log(rmaxf(primary[idx - 1], 1.0e-12f));
                const float center = std::log(rmaxf(primary[idx], 1.0e-12f));
                const float right = std::log(rmaxf(primary[idx + 1], 1.0e-12f));
                const float denom = left - 2.0f * center + right;
                if (std::fabs(denom) > 1.0e-6f) offset = rclampf(0.5f * (left - right) / denom, -0.5f, 0.5f);
            }
            const float spectral = rclampf(best_f + offset * 1.0f, 49.0f, 61.0f);
            const float window_seconds = static_cast<float>(window_samples_) / rmaxf(fs_, 1.0f);
            const float center_sample = (static_cast<float>(windows_observed_) + 0.5f) * static_cast<float>(window_samples_);
            const float absolute_phase =
                wrap_phase(-best_phase + 2.0f * kPiF * best_f * center_sample / rmaxf(fs_, 1.0f));
            bool have_phase = false;
            float phase_hz = 0.0f;
            if (phase_valid_ && window_seconds > 0.0f) {
                const float delta = wrap_phase(absolute_phase - prev_phase_);
                const float base = delta / (2.0f * kPiF * window_seconds);
                const float spacing = 1.0f / window_seconds;
                float best_alias = base, best_err = std::numeric_limits<float>::infinity();
                for (int k = -32; k <= 32; ++k) {
                    const float cand = base + static_cast<float>(k) * spacing;
                    const float err = std::fabs(cand - spectral);
                    if (err < best_err) {
                        best_alias = cand;
                        best_err = err;
                    }
                }
                phase_hz = rclampf(best_alias, 49.0f, 61.0f);
                have_phase = true;
            }
            const float measured = have_phase ? 0.75f * spectral + 0.25f * phase_hz : spectral;
            const float next = hum_line_hz_ <= 0.0f ? measured : hum_line_hz_ + 0.35f * (measured - hum_line_hz_);
            hum_line_hz_ = rclampf(next, 49.0f, 61.0f);
            prev_phase_ = absolute_phase;
            phase_valid_ = true;
        }
    }

    float fs_;
    CleanupMode mode_ = CleanupMode::Off;
    float lowpass_ = 0, low_env_ = 0, slow_low_env_ = 0, broadband_env_ = 0;
    uint32_t rumble_hold_ = 0;
    std::vector<HumBin> bins_, harm_bins_;
    size_t window_samples_ = 1, window_pos_ = 0;
    uint32_t windows_observed_ = 0;
    uint8_t candidate_windows_ = 0;
    float total_energy_ = 0;
    uint32_t hum_hold_ = 0;
    float hum_line_hz_ = 0, prev_phase_ = 0;
    bool phase_valid_ = false;
    float hum_strength_ = 0, harmonic_strength_ = 0;
    Biquad highpass_;
    float highpass_hz_ = 80.0f;
    SmoothNotch hum_notch_, harmonic_notch_;
    bool hum_detected_ = false, rumble_detected_ = false;
    float selected_hp_hz_ = 80.0f;
};

// Input stage in front of the offline chain (new optional key; AF_INPUT_NONE = reference behaviour).
// Block contract of processor/tests.rs:500-549 with 480-sample blocks.
class InputStage {
  public:
    InputStage(int mode, double fs)
        : mode_(mode), hp_(BiquadType::HighPass, 80.0, 0.0, 0.707, fs), cleanup_(static_cast<float>(fs)) {
        if (mode == 2) cleanup_.set_mode(CleanupMode::Gentle);
        if (mode == 3) cleanup_.set_mode(CleanupMode::Strong);
    }
    void process(float* buf, size_t n) {
        if (mode_ == 0) return;
        for (size_t off = 0; off < n; off += 480) {
            const size_t len = std::min<size_t>(480, n - off);
            float* blk = buf + off;
            if (cleanup_.enabled()) cleanup_.analyze_input(blk, len);
            apply_input_pre_filter(blk, len, dc_, hp_, !cleanup_.enabled());
            if (cleanup_.enabled()) cleanup_.process_block(blk, len);
        }
    }
    const AdaptiveInputCleanup& cleanup() const { return cleanup_; }

  private:
    int mode_;
    InputPreFilterState dc_;
    Biquad hp_;
    AdaptiveInputCleanup cleanup_;
};

// ---- audio/processor/block_processor.rs -------------------------------------------------------------
struct OfflineDspBlockStats {
    float input_sample_peak = 0, output_sample_peak = 0, true_peak_limiter_input_peak = 0, output_true_peak = 0;
    float limiter_peak_gain_reduction_db = 0, true_peak_limiter_gain_reduction_db = 0;
    uint64_t true_peak_limited_events = 0;
    float compressor_gain_reduction_db = 0, deesser_gain_reduction_db = 0;
};

class OfflineDspBlockProcessor {
  public:
    explicit OfflineDspBlockProcessor(double fs)  // block_processor.rs:46-60
        : deesser(fs), eq(fs), compressor(-18.0, 3.0, 5.0, 100.0, 0.0, 6.0, fs), limiter(-0.5, 50.0, fs),
          true_peak_limiter(static_cast<float>(fs), -1.5f, 80.0f) {}
    void set_deesser_enabled(bool e) { deesser_enabled_ = e; deesser.set_enabled(e); }
    void set_eq_enabled(bool e) { eq_enabled_ = e; eq.set_enabled(e); }
    void set_compressor_enabled(bool e) { compressor_enabled_ = e; compressor.set_enabled(e); }
    void set_limiter_enabled(bool e) { limiter_enabled_ = e; limiter.set_enabled(e); }
    void set_eq_before_deesser(bool e) { eq_before_deesser_ = e; }

    // block_processor.rs:106-161; `block` is processed in place (input == output buffer here).
    OfflineDspBlockStats process_block_with_stats(float* block, size_t n) {
        OfflineDspBlockStats st;
        for (size_t i = 0; i < n; ++i) st.input_sample_peak = rmaxf(st.input_sample_peak, std::fabs(block[i]));
        if (eq_before_deesser_) {
            if (eq_enabled_) eq.process_block_inplace(block, n);
            if (deesser_enabled_) {
                deesser.process_block_inplace(block, n);
                st.deesser_gain_reduction_db = deesser.current_gain_reduction_db();
            }
        } else {
            if (deesser_enabled_) {
                deesser.process_block_inplace(block, n);
                st.deesser_gain_reduction_db = deesser.current_gain_reduction_db();
            }
            if (eq_enabled_) eq.process_block_inplace(block, n);
        }
        if (compressor_enabled_) {
            compressor.process_block_inplace(block, n);
            st.compressor_gain_reduction_db = static_cast<float>(compressor.current_gain_reduction());
        }
        if (limiter_enabled_) {
            limiter.process_block_inplace(block, n);
            st.limiter_peak_gain_reduction_db = static_cast<float>(limiter.peak_gain_reduction_and_reset());
            true_peak_limiter.set_ceiling_linear(std::pow(10.0f, static_cast<float>(limiter.ceiling_db()) / 20.0f));
            const TruePeakLimiterBlockStats tp = true_peak_limiter.process_block_inplace(block, n);
            st.true_peak_limiter_input_peak = tp.input_true_peak;
            st.true_peak_limiter_gain_reduction_db = tp.max_gain_reduction_db;
            st.true_peak_limited_events = tp.limited_events;
        }
        for (size_t i = 0; i < n; ++i) st.output_sample_peak = rmaxf(st.output_sample_peak, std::fabs(block[i]));
        st.output_true_peak = true_peak_detector.process_block(block, n);
        return st;
    }

    DeEsser deesser;
    ParametricEQ eq;
    Compressor compressor;
    Limiter limiter;
    TruePeakLimiter true_peak_limiter;
    TruePeakDetector true_peak_detector;

  private:
    bool deesser_enabled_ = false, eq_enabled_ = true, compressor_enabled_ = false, limiter_enabled_ = true,
         eq_before_deesser_ = false;
};

// ---- audio/processor/python_api.rs:54-111 -------------------------------------------------------------
inline float linear_to_db_f32(float v) { return 20.0f * std::log10(rmaxf(v, 1.0e-12f)); }

inline bool total_less(float a, float b) {  // f32::total_cmp
    int32_t ia, ib;
    std::memcpy(&ia, &a, 4);
    std::memcpy(&ib, &b, 4);
    ia ^= static_cast<int32_t>(static_cast<uint32_t>(ia >> 31) >> 1);
    ib ^= static_cast<int32_t>(static_cast<uint32_t>(ib >> 31) >> 1);
    return ia < ib;
}

inline float percentile_f32(std::vector<float> values, float percentile) {  // :58-72 (takes a copy; callers clone)
    if (values.empty()) return 0.0f;
    std::stable_sort(values.begin(), values.end(), total_less);
    const float position = static_cast<float>(values.size() - 1) * rclampf(percentile, 0.0f, 1.0f);
    const size_t lower = as_usize(static_cast<double>(std::floor(position)));
    const size_t upper = as_usize(static_cast<double>(std::ceil(position)));
    if (lower == upper) return values[lower];
    const float fraction = position - static_cast<float>(lower);
    return values[lower] + fraction * (values[upper] - values[lower]);
}

inline float compressor_pumping_score(const std::vector<float>& trace, float cadence_hz) {  // :74-111
    if (trace.size() < 3 || !std::isfinite(cadence_hz) || cadence_hz <= 0.0f) return 0.0f;
    const float dt = 1.0f / cadence_hz;
    const float hp_rc = 1.0f / (2.0f * kPiF * 2.0f);
    const float lp_rc = 1.0f / (2.0f * kPiF * 8.0f);
    const float hp_alpha = hp_rc / (hp_rc + dt);
    const float lp_alpha = dt / (lp_rc + dt);
    float prev = trace[0], hp = 0.0f, bp = 0.0f;
    std::vector<float> bp_abs, deltas;
    for (size_t i = 1; i < trace.size(); ++i) {
        const float v = trace[i];
        if (!std::isfinite(v)) return std::numeric_limits<float>::infinity();
        hp = hp_alpha * (hp + v - prev);
        bp += lp_alpha * (hp - bp);
        bp_abs.push_back(std::fabs(bp));
        deltas.push_back(std::fabs(v - prev));
        prev = v;
    }
    const float limit = percentile_f32(bp_abs, 0.95f);
    float robust = 0.0f;
    if (!bp_abs.empty()) {
        float sum = 0.0f;
        for (float v : bp_abs) {
            const float m = rminf(v, limit);
            sum += m * m;
        }
        robust = std::sqrt(sum / static_cast<float>(bp_abs.size()));
    }
    return robust + percentile_f32(deltas, 0.95f);
}

}  // namespace orc
