// afsim_oracle_capi.cpp -- C entry points of the parity oracle (liboracle.so).
//
// TEST INFRASTRUCTURE ONLY (see afsim_oracle.hpp).  Two layers:
//   orc_chain_render / orc_eq_render / orc_eq_response / orc_auto_makeup_control
//       restate the reference's pyfunctions (python_api.rs:118-276,378-714, lib.rs:99-288);
//   orc_proc_* / orc_biquad_* / orc_limiter_* / ...
//       expose the DSP objects one setter at a time, so tests/ can restate the reference's
//       Rust unit tests (e.g. the golden vector processor/tests.rs:1784-1885) verbatim.
// POD structs come from include/afsim.h (the boundary definition; no algorithm is shared).
#include <chrono>
#include <cstdio>
#include <string>
#include <thread>

#include "../include/afsim.h"
#include "afsim_oracle.hpp"

using namespace orc;

namespace {
thread_local std::string g_error;
int fail(const std::string& msg) {
    g_error = msg;
    return AFSIM_INVALID_ARGUMENT;
}

EqBandConfig typed_config(const AfBand& b) {
    return {static_cast<EqFilterType>(b.filter_type), b.frequency_hz, b.gain_db, b.q, b.slope_db_per_octave, b.enabled != 0};
}

// lib.rs:154-189 parse_eq_v2_bands
int parse_typed(const AfBand* bands, double fs, std::vector<EqBandConfig>& out) {
    if (!std::isfinite(fs) || fs <= 0.0) return fail("sample_rate must be finite and positive");
    for (size_t i = 0; i < 10; ++i) {
        if (bands[i].filter_type > 5)
            return fail("band " + std::to_string(i) + " has unsupported EQ filter type: " + std::to_string(bands[i].filter_type));
        EqBandConfig c = typed_config(bands[i]);
        const std::string msg = validate_band(c, i, fs);
        if (!msg.empty()) return fail(msg);
        out.push_back(c);
    }
    return AFSIM_OK;
}

// processor/control.rs:904-910
double effective_limiter_ceiling_db(double ceiling_db, bool careful) { return careful ? rmin(ceiling_db, -1.5) : ceiling_db; }

// python_api.rs:400-487: constructor + setter order of simulate_auto_eq_chain
void configure(OfflineDspBlockProcessor& p, const AfBand* bands, const std::vector<EqBandConfig>* typed,
               const AfChainSettings& s, float& effective_ceiling_db) {
    p.set_eq_enabled(true);
    if (typed) {
        for (size_t i = 0; i < 10; ++i) p.eq.set_band_config(i, (*typed)[i]);
        p.eq.reset();
    } else {
        for (size_t i = 0; i < 10; ++i) {
            p.eq.set_band_frequency(i, bands[i].frequency_hz);
            p.eq.set_band_gain(i, bands[i].gain_db);
            p.eq.set_band_q(i, bands[i].q);
        }
    }
    p.set_eq_before_deesser(s.eq_before_deesser != 0);
    p.set_deesser_enabled(s.deesser_enabled != 0);
    if (s.deesser_enabled) {
        p.deesser.set_auto_enabled(s.deesser_auto_enabled != 0);
        p.deesser.set_auto_amount(s.deesser_auto_amount);
        p.deesser.set_low_cut_hz(s.deesser_low_cut_hz);
        p.deesser.set_high_cut_hz(s.deesser_high_cut_hz);
        p.deesser.set_threshold_db(s.deesser_threshold_db);
        p.deesser.set_ratio(s.deesser_ratio);
        p.deesser.set_attack_ms(s.deesser_attack_ms);
        p.deesser.set_release_ms(s.deesser_release_ms);
        p.deesser.set_max_reduction_db(s.deesser_max_reduction_db);
    }
    p.set_compressor_enabled(s.compressor_enabled != 0);
    if (s.compressor_enabled) {
        p.compressor.set_threshold(s.compressor_threshold_db);
        p.compressor.set_ratio(s.compressor_ratio);
        p.compressor.set_attack_time(s.compressor_attack_ms);
        p.compressor.set_release_time(s.compressor_release_ms);
        p.compressor.set_makeup_gain(s.compressor_makeup_gain_db);
        p.compressor.set_adaptive_release(s.compressor_adaptive_release != 0);
        p.compressor.set_base_release_time(s.compressor_base_release_ms);
        p.compressor.set_auto_makeup_enabled(s.compressor_auto_makeup_enabled != 0);
        p.compressor.set_target_lufs(s.compressor_target_lufs);
        p.compressor.set_sidechain_highpass_enabled(s.compressor_sidechain_highpass_enabled != 0);
    }
    p.set_limiter_enabled(s.limiter_enabled != 0);
    effective_ceiling_db =
        static_cast<float>(effective_limiter_ceiling_db(s.limiter_ceiling_db, s.limiter_careful_output_enabled != 0));
    if (s.limiter_enabled) {
        p.limiter.set_lookahead_ms(s.limiter_lookahead_ms);
        p.limiter.set_ceiling(static_cast<double>(effective_ceiling_db));
        p.limiter.set_release_time(s.limiter_release_ms);
        p.true_peak_limiter.set_release_ms(static_cast<float>(s.limiter_release_ms));
    }
}
}  // namespace

extern "C" {

const char* orc_last_error() { return g_error.c_str(); }

void orc_chain_settings_default(AfChainSettings* s) {  // python_api.rs:415-487 defaults
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

// simulate_auto_eq_chain, python_api.rs:378-714.  rows_out (nullable): 4 floats per analysis block.
int orc_chain_render(const float* audio_in, size_t n, double fs, const AfBand* bands, const AfChainSettings* s,
                     AfChainMetrics* m, float* out_audio, float* rows_out, size_t rows_capacity) {
    const auto started = std::chrono::steady_clock::now();
    if (!std::isfinite(fs) || fs <= 0.0) return fail("sample_rate must be positive and finite");
    std::vector<EqBandConfig> typed;
    if (s->use_typed_bands) {
        const int rc = parse_typed(bands, fs, typed);
        if (rc != AFSIM_OK) return rc;
    }
    OfflineDspBlockProcessor proc(fs);
    float effective_ceiling_db = 0.0f;
    configure(proc, bands, s->use_typed_bands ? &typed : nullptr, *s, effective_ceiling_db);

    // Optional input stage (new key; absent in the reference's offline simulator).
    std::vector<float> staged;
    const float* audio = audio_in;
    if (s->input_stage != AF_INPUT_NONE) {
        staged.assign(audio_in, audio_in + n);
        for (float& v : staged)
            if (!std::isfinite(v)) v = 0.0f;
        InputStage stage(s->input_stage, fs);
        stage.process(staged.data(), n);
        audio = staged.data();
    }

    double input_square_sum = 0.0, output_square_sum = 0.0;
    size_t input_samples = 0, output_samples = 0;
    float input_sample_peak = 0, output_sample_peak = 0, pre_limiter_true_peak = 0, output_true_peak = 0;
    float limiter_gr = 0, tp_gr = 0, comp_gr = 0, de_gr = 0;
    uint64_t limited_events = 0;
    struct Row { float in_db, out_db, comp, de; };
    std::vector<Row> rows;
    bool non_finite_output = false;

    const size_t block_samples = std::clamp<size_t>(as_usize(std::round(fs * 0.020)), 1, 8192);
    std::vector<float> block;
    for (size_t off = 0; off < n; off += block_samples) {
        const size_t len = std::min(block_samples, n - off);
        block.assign(audio + off, audio + off + len);
        double block_in_sq = 0.0;
        for (float& v : block) {
            if (!std::isfinite(v)) v = 0.0f;
            input_square_sum += static_cast<double>(v) * static_cast<double>(v);
            block_in_sq += static_cast<double>(v) * static_cast<double>(v);
            input_samples += 1;
        }
        const OfflineDspBlockStats st = proc.process_block_with_stats(block.data(), len);
        const float in_rms = static_cast<float>(std::sqrt(block_in_sq / static_cast<double>(len)));
        double block_out_sq = 0.0;
        for (float v : block) {
            if (!std::isfinite(v)) {
                non_finite_output = true;
            } else {
                block_out_sq += static_cast<double>(v) * static_cast<double>(v);
            }
        }
        const float out_rms = static_cast<float>(std::sqrt(block_out_sq / static_cast<double>(len)));
        rows.push_back({linear_to_db_f32(in_rms), linear_to_db_f32(out_rms), st.compressor_gain_reduction_db,
                        st.deesser_gain_reduction_db});
        input_sample_peak = rmaxf(input_sample_peak, st.input_sample_peak);
        output_sample_peak = rmaxf(output_sample_peak, st.output_sample_peak);
        pre_limiter_true_peak = rmaxf(pre_limiter_true_peak, st.true_peak_limiter_input_peak);
        output_true_peak = rmaxf(output_true_peak, st.output_true_peak);
        limiter_gr = rmaxf(limiter_gr, st.limiter_peak_gain_reduction_db);
        tp_gr = rmaxf(tp_gr, st.true_peak_limiter_gain_reduction_db);
        comp_gr = rmaxf(comp_gr, st.compressor_gain_reduction_db);
        de_gr = rmaxf(de_gr, st.deesser_gain_reduction_db);
        limited_events += st.true_peak_limited_events;
        for (float v : block) {
            output_square_sum += static_cast<double>(v) * static_cast<double>(v);
            output_samples += 1;
        }
        if (out_audio) std::memcpy(out_audio + off, block.data(), len * sizeof(float));
    }

    // python_api.rs:578-648 reductions
    const float input_rms = input_samples ? static_cast<float>(std::sqrt(input_square_sum / static_cast<double>(input_samples))) : 0.0f;
    const float output_rms = output_samples ? static_cast<float>(std::sqrt(output_square_sum / static_cast<double>(output_samples))) : 0.0f;
    const float output_sample_peak_db = linear_to_db_f32(output_sample_peak);
    const float pre_limiter_true_peak_db = linear_to_db_f32(pre_limiter_true_peak);
    const float output_true_peak_db = linear_to_db_f32(output_true_peak);
    std::vector<float> in_rows;
    for (const Row& r : rows) in_rows.push_back(r.in_db);
    const float floor_db = percentile_f32(in_rows, 0.20f);
    const float p90_db = percentile_f32(in_rows, 0.90f);
    const float active_thr = rmaxf(rmaxf(floor_db + 6.0f, p90_db - 24.0f), -60.0f);
    std::vector<float> act_comp, act_de;
    for (const Row& r : rows)
        if (r.in_db >= active_thr) {
            act_comp.push_back(rmaxf(r.comp, 0.0f));
            act_de.push_back(rmaxf(r.de, 0.0f));
        }
    if (act_comp.size() < 3) {
        act_comp.clear();
        act_de.clear();
        for (const Row& r : rows) {
            act_comp.push_back(rmaxf(r.comp, 0.0f));
            act_de.push_back(rmaxf(r.de, 0.0f));
        }
    }
    const size_t active_count = act_comp.size();
    float active_ratio = 0.0f;
    if (active_count > 0) {
        size_t c = 0;
        for (float v : act_comp) c += v >= 0.10f ? 1 : 0;
        active_ratio = static_cast<float>(c) / static_cast<float>(active_count);
    }
    std::vector<float> act_gain, sil_delta, sil_gain, gr_trace;
    for (const Row& r : rows) {
        if (r.in_db >= active_thr && r.in_db > -100.0f) act_gain.push_back(r.out_db - r.in_db);
        if (r.in_db < active_thr && r.in_db > -100.0f) sil_delta.push_back(r.out_db - r.in_db);
        if (r.in_db < active_thr) sil_gain.push_back(-rmaxf(r.comp, 0.0f));
        gr_trace.push_back(rmaxf(r.comp, 0.0f));
    }

    std::memset(m, 0, sizeof *m);
    m->input_sample_peak_db = linear_to_db_f32(input_sample_peak);
    m->input_rms_db = linear_to_db_f32(input_rms);
    m->output_sample_peak_db = output_sample_peak_db;
    m->pre_limiter_true_peak_db = pre_limiter_true_peak_db;
    m->output_true_peak_db = output_true_peak_db;
    m->output_rms_db = linear_to_db_f32(output_rms);
    m->limiter_effective_ceiling_db = effective_ceiling_db;
    m->sample_headroom_db = effective_ceiling_db - output_sample_peak_db;
    m->pre_limiter_true_peak_headroom_db = effective_ceiling_db - pre_limiter_true_peak_db;
    m->true_peak_headroom_db = effective_ceiling_db - output_true_peak_db;
    m->limiter_gain_reduction_db = limiter_gr;
    m->true_peak_limiter_gain_reduction_db = tp_gr;
    m->true_peak_limited_events = limited_events;
    m->compressor_gain_reduction_db = comp_gr;
    m->deesser_gain_reduction_db = de_gr;
    m->compressor_gain_reduction_median_db = percentile_f32(act_comp, 0.50f);
    m->compressor_gain_reduction_p95_db = percentile_f32(act_comp, 0.95f);
    m->compressor_gain_reduction_active_ratio = active_ratio;
    m->active_output_gain_db = percentile_f32(act_gain, 0.50f);
    m->silence_output_gain_db = percentile_f32(sil_gain, 0.50f);
    m->silence_level_delta_db = percentile_f32(sil_delta, 0.50f);
    m->compressor_pumping_score_db = compressor_pumping_score(gr_trace, 50.0f);
    m->non_finite_output = non_finite_output ? 1 : 0;
    m->deesser_gain_reduction_median_db = percentile_f32(act_de, 0.50f);
    m->deesser_gain_reduction_p95_db = percentile_f32(act_de, 0.95f);
    m->analysis_block_ms = 20.0f;
    m->active_analysis_threshold_db = active_thr;
    m->active_analysis_block_count = active_count;
    m->processed_samples = output_samples;
    if (rows_out) {
        const size_t k = std::min(rows.size(), rows_capacity);
        for (size_t i = 0; i < k; ++i) {
            rows_out[4 * i + 0] = rows[i].in_db;
            rows_out[4 * i + 1] = rows[i].out_db;
            rows_out[4 * i + 2] = rows[i].comp;
            rows_out[4 * i + 3] = rows[i].de;
        }
    }
    m->candidate_runtime_ms =
        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - started).count();
    return AFSIM_OK;
}

// Multi-threaded driver over independent streams (what a parallel CPU user of the reference
// would run; used only as bench.py's CPU baseline).  pair i = (passage pair_passage[i], candidate
// pair_candidate[i]).
int orc_chain_sweep(const float* const* passages, const size_t* passage_len, double fs, const AfCandidate* candidates,
                    const uint32_t* pair_passage, const uint32_t* pair_candidate, size_t n_pairs, AfChainMetrics* out,
                    int n_threads) {
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> pool;
    std::vector<int> rcs(static_cast<size_t>(n_threads), AFSIM_OK);
    for (int t = 0; t < n_threads; ++t) {
        pool.emplace_back([&, t]() {
            for (size_t i = static_cast<size_t>(t); i < n_pairs; i += static_cast<size_t>(n_threads)) {
                const AfCandidate& c = candidates[pair_candidate[i]];
                const int rc = orc_chain_render(passages[pair_passage[i]], passage_len[pair_passage[i]], fs, c.bands,
                                                &c.settings, &out[i], nullptr, nullptr, 0);
                if (rc != AFSIM_OK) rcs[static_cast<size_t>(t)] = rc;
            }
        });
    }
    for (auto& th : pool) th.join();
    for (int rc : rcs)
        if (rc != AFSIM_OK) return rc;
    return AFSIM_OK;
}

// simulate_eq_v2, lib.rs:214-288
int orc_eq_render(const float* audio, size_t n, double fs, const AfBand* bands, AfEqRenderStats* st, float* out_audio) {
    std::vector<EqBandConfig> typed;
    const int rc = parse_typed(bands, fs, typed);
    if (rc != AFSIM_OK) return rc;
    for (size_t i = 0; i < n; ++i)
        if (!std::isfinite(audio[i])) return fail("audio must contain only finite samples");
    ParametricEQ eq(fs);
    for (size_t i = 0; i < 10; ++i) eq.set_band_config(i, typed[i]);
    eq.reset();
    std::vector<float> out(audio, audio + n);
    const auto started = std::chrono::steady_clock::now();
    eq.process_block_inplace(out.data(), n);
    const double runtime_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - started).count();
    double in_sq = 0.0, out_sq = 0.0;
    float in_peak = 0.0f, out_peak = 0.0f;
    bool non_finite = false;
    for (size_t i = 0; i < n; ++i) {
        in_sq += static_cast<double>(audio[i]) * static_cast<double>(audio[i]);
        out_sq += static_cast<double>(out[i]) * static_cast<double>(out[i]);
        in_peak = rmaxf(in_peak, std::fabs(audio[i]));
        out_peak = rmaxf(out_peak, std::fabs(out[i]));
        if (!std::isfinite(out[i])) non_finite = true;
    }
    const double divisor = static_cast<double>(std::max<size_t>(n, 1));
    TruePeakDetector din, dout;
    std::vector<double> freqs(512);
    for (int i = 0; i < 512; ++i) freqs[i] = 20.0 * std::pow(20000.0 / 20.0, static_cast<double>(i) / 511.0);
    const std::vector<double> resp = eq.magnitude_response_db(freqs.data(), freqs.size());
    double max_resp = -std::numeric_limits<double>::infinity();
    for (double v : resp) max_resp = rmax(max_resp, v);
    std::memset(st, 0, sizeof *st);
    st->input_sample_peak = in_peak;
    st->output_sample_peak = out_peak;
    st->input_true_peak = din.process_block(audio, n);
    st->output_true_peak = dout.process_block(out.data(), n);
    st->input_rms = std::sqrt(in_sq / divisor);
    st->output_rms = std::sqrt(out_sq / divisor);
    st->max_response_db = max_resp;
    st->runtime_ms = runtime_ms;
    st->sample_count = n;
    st->algorithmic_latency_samples = 0;
    st->non_finite_output = non_finite ? 1 : 0;
    if (out_audio) std::memcpy(out_audio, out.data(), n * sizeof(float));
    return AFSIM_OK;
}

// eq_magnitude_response (typed=0, lib.rs:99-150) / eq_magnitude_response_v2 (typed=1, lib.rs:191-212)
int orc_eq_response(const double* freqs, size_t n_freqs, const AfBand* bands, int typed, double fs, double* out_db) {
    ParametricEQ eq(fs > 0.0 && std::isfinite(fs) ? fs : 48000.0);
    if (typed) {
        std::vector<EqBandConfig> cfg;
        const int rc = parse_typed(bands, fs, cfg);
        if (rc != AFSIM_OK) return rc;
        const double nyquist = fs / 2.0;
        for (size_t i = 0; i < n_freqs; ++i)
            if (!std::isfinite(freqs[i]) || freqs[i] < 0.0 || freqs[i] > nyquist)
                return fail("response frequencies must be finite and between 0 Hz and Nyquist");
        for (size_t i = 0; i < 10; ++i) eq.set_band_config(i, cfg[i]);
    } else {
        if (!std::isfinite(fs) || fs <= 0.0) return fail("sample_rate must be finite and positive");
        const double nyquist = fs / 2.0;
        for (size_t i = 0; i < 10; ++i) {
            const AfBand& b = bands[i];
            if (!std::isfinite(b.frequency_hz) || b.frequency_hz <= 0.0 || b.frequency_hz >= nyquist)
                return fail("band " + std::to_string(i) + " frequency must be between 0 Hz and Nyquist");
            if (!std::isfinite(b.gain_db)) return fail("band " + std::to_string(i) + " gain must be finite");
            if (!std::isfinite(b.q) || b.q <= 0.0) return fail("band " + std::to_string(i) + " Q must be finite and positive");
        }
        for (size_t i = 0; i < n_freqs; ++i)
            if (!std::isfinite(freqs[i]) || freqs[i] < 0.0 || freqs[i] > nyquist)
                return fail("response frequencies must be finite and between 0 Hz and Nyquist");
        for (size_t i = 0; i < 10; ++i) {
            eq.set_band_frequency(i, bands[i].frequency_hz);
            eq.set_band_gain(i, bands[i].gain_db);
            eq.set_band_q(i, bands[i].q);
        }
    }
    const std::vector<double> r = eq.magnitude_response_db(freqs, n_freqs);
    std::memcpy(out_db, r.data(), n_freqs * sizeof(double));
    return AFSIM_OK;
}

// simulate_auto_makeup_control, python_api.rs:118-276.  traces: 6 arrays of block_count floats
// (makeup_gain_db, activity, reliability, gain_reduction_db, input_rms_db, output_rms_db).
int orc_auto_makeup_control(const float* audio, size_t n, double fs, const double* vad, size_t n_vad, double noise_floor_db,
                            double noise_reliability, double threshold_db, double ratio, double attack_ms,
                            double release_ms, double makeup_gain_db, double target_lufs, int adaptive_release,
                            int sidechain_hp, double vad_reliability, float* traces, float* out_audio) {
    if (!std::isfinite(fs) || fs <= 0.0) return fail("sample_rate must be positive and finite");
    if (!std::isfinite(noise_floor_db) || !std::isfinite(noise_reliability) || !(noise_reliability >= 0.0 && noise_reliability <= 1.0))
        return fail("noise evidence must be finite and reliability must be between 0 and 1");
    for (size_t i = 0; i < n_vad; ++i)
        if (!std::isfinite(vad[i]) || !(vad[i] >= 0.0 && vad[i] <= 1.0))
            return fail("VAD probabilities must be finite and between 0 and 1");
    const size_t block_count = (n + 479) / 480;
    if (n_vad != 0 && n_vad != block_count)
        return fail("expected " + std::to_string(block_count) + " VAD probabilities at the 10 ms control cadence, got " +
                    std::to_string(n_vad));
    Compressor comp(threshold_db, ratio, attack_ms, release_ms, makeup_gain_db, 6.0, fs);
    comp.set_auto_makeup_enabled(true);
    comp.set_target_lufs(target_lufs);
    comp.set_noise_reference_reliability(noise_reliability);
    comp.set_adaptive_release(adaptive_release != 0);
    comp.set_sidechain_highpass_enabled(sidechain_hp != 0);
    if (!std::isfinite(vad_reliability) || !(vad_reliability >= 0.0 && vad_reliability <= 1.0))
        return fail("vad_reliability must be finite and between 0 and 1");
    std::vector<float> block;
    for (size_t bi = 0; bi < block_count; ++bi) {
        const size_t off = bi * 480, len = std::min<size_t>(480, n - off);
        block.assign(audio + off, audio + off + len);
        double sq = 0.0;
        for (float v : block) sq += static_cast<double>(v) * static_cast<double>(v);
        const float in_rms = static_cast<float>(std::sqrt(sq / static_cast<double>(std::max<size_t>(len, 1))));
        AutoMakeupActivityInput ev{0, vad_reliability, noise_floor_db, noise_reliability};
        const AutoMakeupActivityInput* evp = nullptr;
        if (bi < n_vad) {
            ev.vad_probability = vad[bi];
            evp = &ev;
        }
        comp.process_block_with_activity(block.data(), len, evp);
        sq = 0.0;
        for (float v : block) sq += static_cast<double>(v) * static_cast<double>(v);
        const float out_rms = static_cast<float>(std::sqrt(sq / static_cast<double>(std::max<size_t>(len, 1))));
        traces[0 * block_count + bi] = static_cast<float>(comp.current_makeup_gain());
        traces[1 * block_count + bi] = static_cast<float>(comp.auto_makeup_activity());
        traces[2 * block_count + bi] = static_cast<float>(comp.auto_makeup_activity_reliability());
        traces[3 * block_count + bi] = static_cast<float>(comp.current_gain_reduction());
        traces[4 * block_count + bi] = linear_to_db_f32(in_rms);
        traces[5 * block_count + bi] = linear_to_db_f32(out_rms);
        if (out_audio) std::memcpy(out_audio + off, block.data(), len * sizeof(float));
    }
    return AFSIM_OK;
}

// ---- object-level API (restating the reference's Rust unit tests from Python) -------------------------

// OfflineDspBlockProcessor
void* orc_proc_new(double fs) { return new OfflineDspBlockProcessor(fs); }
void orc_proc_free(void* p) { delete static_cast<OfflineDspBlockProcessor*>(p); }
#define PROC static_cast<OfflineDspBlockProcessor*>(p)
void orc_proc_set_deesser_enabled(void* p, int e) { PROC->set_deesser_enabled(e != 0); }
void orc_proc_set_eq_enabled(void* p, int e) { PROC->set_eq_enabled(e != 0); }
void orc_proc_set_compressor_enabled(void* p, int e) { PROC->set_compressor_enabled(e != 0); }
void orc_proc_set_limiter_enabled(void* p, int e) { PROC->set_limiter_enabled(e != 0); }
void orc_proc_set_eq_before_deesser(void* p, int e) { PROC->set_eq_before_deesser(e != 0); }
void orc_proc_deesser_set_auto_enabled(void* p, int e) { PROC->deesser.set_auto_enabled(e != 0); }
void orc_proc_deesser_set_auto_amount(void* p, double v) { PROC->deesser.set_auto_amount(v); }
void orc_proc_deesser_set_max_reduction_db(void* p, double v) { PROC->deesser.set_max_reduction_db(v); }
void orc_proc_eq_set_band_frequency(void* p, size_t i, double v) { PROC->eq.set_band_frequency(i, v); }
void orc_proc_eq_set_band_gain(void* p, size_t i, double v) { PROC->eq.set_band_gain(i, v); }
void orc_proc_eq_set_band_q(void* p, size_t i, double v) { PROC->eq.set_band_q(i, v); }
void orc_proc_comp_set_threshold(void* p, double v) { PROC->compressor.set_threshold(v); }
void orc_proc_comp_set_ratio(void* p, double v) { PROC->compressor.set_ratio(v); }
void orc_proc_comp_set_attack_time(void* p, double v) { PROC->compressor.set_attack_time(v); }
void orc_proc_comp_set_release_time(void* p, double v) { PROC->compressor.set_release_time(v); }
void orc_proc_comp_set_makeup_gain(void* p, double v) { PROC->compressor.set_makeup_gain(v); }
void orc_proc_comp_set_adaptive_release(void* p, int e) { PROC->compressor.set_adaptive_release(e != 0); }
void orc_proc_limiter_set_ceiling(void* p, double v) { PROC->limiter.set_ceiling(v); }
void orc_proc_limiter_set_release_time(void* p, double v) { PROC->limiter.set_release_time(v); }
// stats: [input_peak, output_peak, tp_in_peak, out_true_peak, lim_gr, tp_gr, comp_gr, de_gr, limited_events]
void orc_proc_process_block(void* p, float* block, size_t n, float* stats9) {
    const OfflineDspBlockStats st = PROC->process_block_with_stats(block, n);
    stats9[0] = st.input_sample_peak;
    stats9[1] = st.output_sample_peak;
    stats9[2] = st.true_peak_limiter_input_peak;
    stats9[3] = st.output_true_peak;
    stats9[4] = st.limiter_peak_gain_reduction_db;
    stats9[5] = st.true_peak_limiter_gain_reduction_db;
    stats9[6] = st.compressor_gain_reduction_db;
    stats9[7] = st.deesser_gain_reduction_db;
    stats9[8] = static_cast<float>(st.true_peak_limited_events);
}
#undef PROC

// Biquad (type ids: 0 LowShelf, 1 HighShelf, 2 Peaking, 3 Notch, 4 HighPass, 5 LowPass, 6 Bypass)
void* orc_biquad_new(int type, double f, double g, double q, double fs) { return new Biquad(static_cast<BiquadType>(type), f, g, q, fs); }
void orc_biquad_free(void* b) { delete static_cast<Biquad*>(b); }
void orc_biquad_process(void* b, float* buf, size_t n) { static_cast<Biquad*>(b)->process_block_inplace(buf, n); }
double orc_biquad_response_db(void* b, double f) { return static_cast<Biquad*>(b)->magnitude_response_db(f); }
void orc_biquad_set_gain_db(void* b, double g) { static_cast<Biquad*>(b)->set_gain_db(g); }
void orc_biquad_set_frequency(void* b, double f) { static_cast<Biquad*>(b)->set_frequency(f); }
void orc_biquad_reset(void* b) { static_cast<Biquad*>(b)->reset(); }
int orc_biquad_is_crossfading(void* b) { return static_cast<Biquad*>(b)->is_crossfading() ? 1 : 0; }
void orc_biquad_coeffs(void* b, double* out5) {
    const Coeffs& c = static_cast<Biquad*>(b)->active();
    out5[0] = c.b0; out5[1] = c.b1; out5[2] = c.b2; out5[3] = c.a1; out5[4] = c.a2;
}

// Compressor
void* orc_comp_new(double thr, double ratio, double attack, double release, double makeup, double knee, double fs) {
    return new Compressor(thr, ratio, attack, release, makeup, knee, fs);
}
void orc_comp_free(void* c) { delete static_cast<Compressor*>(c); }
double orc_comp_compute_gain_reduction(void* c, double db) { return static_cast<Compressor*>(c)->compute_gain_reduction(db); }
double orc_comp_blended_detector_db(double peak_db, double rms_db) { return Compressor::blended_detector_db(peak_db, rms_db); }
void orc_comp_set_adaptive_release(void* c, int e) { static_cast<Compressor*>(c)->set_adaptive_release(e != 0); }
void orc_comp_set_sidechain_highpass_enabled(void* c, int e) { static_cast<Compressor*>(c)->set_sidechain_highpass_enabled(e != 0); }
void orc_comp_set_auto_makeup_enabled(void* c, int e) { static_cast<Compressor*>(c)->set_auto_makeup_enabled(e != 0); }
void orc_comp_process_block(void* c, float* buf, size_t n) { static_cast<Compressor*>(c)->process_block_inplace(buf, n); }
void orc_comp_process_samples(void* c, float* buf, size_t n) {
    for (size_t i = 0; i < n; ++i) buf[i] = static_cast<Compressor*>(c)->process_sample(buf[i]);
}
void orc_comp_set_target_lufs(void* c, double t) { static_cast<Compressor*>(c)->set_target_lufs(t); }
void orc_comp_set_noise_reference_reliability(void* c, double r) { static_cast<Compressor*>(c)->set_noise_reference_reliability(r); }
// process_block_inplace_with_activity_control (compressor.rs:700-722): evidence = Some(..) when has_evidence
void orc_comp_process_block_with_activity(void* c, float* buf, size_t n, int has_evidence, double vad_probability,
                                          double vad_reliability, double noise_floor_db, double live_noise_reliability) {
    const AutoMakeupActivityInput ev{vad_probability, vad_reliability, noise_floor_db, live_noise_reliability};
    static_cast<Compressor*>(c)->process_block_with_activity(buf, n, has_evidence ? &ev : nullptr);
}
// dsp/loudness.rs: LoudnessMeter::new fails for a rate outside the list (:35-40) -> nullptr
void* orc_meter_new(uint32_t fs) { return LoudnessMeter::supported_rate(fs) ? new LoudnessMeter(fs) : nullptr; }
void orc_meter_free(void* m) { delete static_cast<LoudnessMeter*>(m); }
void orc_meter_process(void* m, const float* x, size_t n) { static_cast<LoudnessMeter*>(m)->process(x, n); }
float orc_meter_momentary(void* m) { return static_cast<LoudnessMeter*>(m)->loudness_momentary(); }
void orc_meter_reset(void* m) { static_cast<LoudnessMeter*>(m)->reset(); }
void orc_comp_update_auto_makeup_gain(void* c, double activity, double reliability, size_t elapsed) {
    static_cast<Compressor*>(c)->test_update_auto_makeup_gain(activity, reliability, elapsed);
}
void orc_comp_estimate_activity(void* c, double rms_db, int has_evidence, double vad_probability, double vad_reliability,
                                double noise_floor_db, double live_noise_reliability, double* out2) {
    const AutoMakeupActivityInput ev{vad_probability, vad_reliability, noise_floor_db, live_noise_reliability};
    static_cast<Compressor*>(c)->test_estimate_activity(rms_db, has_evidence ? &ev : nullptr, &out2[0], &out2[1]);
}
double orc_comp_speech_activity_from_rms_db(double rms_db) { return Compressor::test_speech_activity_from_rms_db(rms_db); }
double orc_comp_auto_makeup_activity(void* c) { return static_cast<Compressor*>(c)->auto_makeup_activity(); }
double orc_comp_auto_makeup_activity_reliability(void* c) { return static_cast<Compressor*>(c)->auto_makeup_activity_reliability(); }
void orc_comp_set_limiter_feedback_gain_reduction_db(void* c, double db) {
    static_cast<Compressor*>(c)->set_limiter_feedback_gain_reduction_db(db);
}
void orc_comp_set_makeup_gain(void* c, double db) { static_cast<Compressor*>(c)->set_makeup_gain(db); }
double orc_comp_gain_reduction(void* c) { return static_cast<Compressor*>(c)->current_gain_reduction(); }
double orc_comp_makeup_gain(void* c) { return static_cast<Compressor*>(c)->current_makeup_gain(); }
double orc_comp_plosive_ratio(void* c) { return static_cast<Compressor*>(c)->plosive_ratio(); }

// Limiter
void* orc_limiter_new(double ceiling_db, double release_ms, double fs, double lookahead_ms) {
    return new Limiter(ceiling_db, release_ms, fs, lookahead_ms);
}
void orc_limiter_free(void* l) { delete static_cast<Limiter*>(l); }
size_t orc_limiter_lookahead_samples(void* l) { return static_cast<Limiter*>(l)->lookahead_samples(); }
void orc_limiter_set_lookahead_ms(void* l, double ms) { static_cast<Limiter*>(l)->set_lookahead_ms(ms); }
void orc_limiter_process(void* l, float* buf, size_t n) { static_cast<Limiter*>(l)->process_block_inplace(buf, n); }
double orc_limiter_peak_gr_and_reset(void* l) { return static_cast<Limiter*>(l)->peak_gain_reduction_and_reset(); }

// True peak
void* orc_tpd_new() { return new TruePeakDetector(); }
void orc_tpd_free(void* d) { delete static_cast<TruePeakDetector*>(d); }
float orc_tpd_process(void* d, const float* buf, size_t n) { return static_cast<TruePeakDetector*>(d)->process_block(buf, n); }
void* orc_tpl_new(float fs, float ceiling_db, float release_ms) { return new TruePeakLimiter(fs, ceiling_db, release_ms); }
void orc_tpl_free(void* t) { delete static_cast<TruePeakLimiter*>(t); }
void orc_tpl_set_ceiling_linear(void* t, float c) { static_cast<TruePeakLimiter*>(t)->set_ceiling_linear(c); }
// stats4: [limited_events, input_true_peak, output_true_peak, max_gain_reduction_db]
void orc_tpl_process(void* t, float* buf, size_t n, float* stats4) {
    const TruePeakLimiterBlockStats st = static_cast<TruePeakLimiter*>(t)->process_block_inplace(buf, n);
    stats4[0] = static_cast<float>(st.limited_events);
    stats4[1] = st.input_true_peak;
    stats4[2] = st.output_true_peak;
    stats4[3] = st.max_gain_reduction_db;
}

// Input stage (mode = AfInputStage); info4: [hum_line_hz, hum_detected, rumble_detected, selected_high_pass_hz]
void orc_input_stage_process(int mode, double fs, float* buf, size_t n, float* info4) {
    InputStage st(mode, fs);
    st.process(buf, n);
    if (info4) {
        info4[0] = st.cleanup().hum_line_hz();
        info4[1] = st.cleanup().hum_detected() ? 1.0f : 0.0f;
        info4[2] = st.cleanup().rumble_detected() ? 1.0f : 0.0f;
        info4[3] = st.cleanup().selected_high_pass_hz();
    }
}
// processor/tests.rs:500-549 `process_adaptive_input_cleanup`: per chunk analyse the raw block, DC block (+ the fixed
// high-pass when the mode is Off), process; info5: [ever_hum, ever_rumble, max selected_high_pass_hz (starting from
// INPUT_PREFILTER_HZ), final hum_line_hz, final selected_high_pass_hz].  mode: 0 off, 2 gentle, 3 strong.
void orc_cleanup_harness(int mode, float fs, float* buf, size_t n, size_t chunk, float* info5) {
    InputPreFilterState dc;
    Biquad hp(BiquadType::HighPass, 80.0, 0.0, 0.707, static_cast<double>(fs));
    AdaptiveInputCleanup cleanup(fs);
    if (mode == 2) cleanup.set_mode(CleanupMode::Gentle);
    if (mode == 3) cleanup.set_mode(CleanupMode::Strong);
    bool ever_hum = false, ever_rumble = false;
    float max_hp = 80.0f;
    for (size_t off = 0; off < n; off += chunk) {
        const size_t len = std::min(chunk, n - off);
        float* blk = buf + off;
        if (cleanup.enabled()) cleanup.analyze_input(blk, len);
        apply_input_pre_filter(blk, len, dc, hp, !cleanup.enabled());
        if (cleanup.enabled()) cleanup.process_block(blk, len);
        ever_hum = ever_hum || cleanup.hum_detected();
        ever_rumble = ever_rumble || cleanup.rumble_detected();
        max_hp = rmaxf(max_hp, cleanup.selected_high_pass_hz());
    }
    info5[0] = ever_hum ? 1.0f : 0.0f;
    info5[1] = ever_rumble ? 1.0f : 0.0f;
    info5[2] = max_hp;
    info5[3] = cleanup.hum_line_hz();
    info5[4] = cleanup.selected_high_pass_hz();
}
// routing.rs:616-641 harness: analyze only; info3: [hum_line_hz, phase_valid, hum_hold_samples]
void orc_cleanup_analyze(int mode, float fs, const float* buf, size_t n, float* info3) {
    AdaptiveInputCleanup c(fs);
    c.set_mode(static_cast<CleanupMode>(mode));
    c.analyze_input(buf, n);
    info3[0] = c.hum_line_hz();
    info3[1] = c.hum_phase_valid() ? 1.0f : 0.0f;
    info3[2] = static_cast<float>(c.hum_hold_samples());
}

float orc_percentile_f32(const float* v, size_t n, float p) { return percentile_f32(std::vector<float>(v, v + n), p); }
float orc_pumping_score(const float* v, size_t n, float cadence) { return compressor_pumping_score(std::vector<float>(v, v + n), cadence); }
double orc_time_constant_to_coeff(double ms, double fs) { return time_constant_to_coeff(ms, fs); }
double orc_butterworth_q(size_t i, size_t n) { return butterworth_section_q(i, n); }

}  // extern "C"
