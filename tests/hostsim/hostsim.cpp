// hostsim.cpp -- TEST HARNESS ONLY: runs the stage-kernel bodies of audio_forge_b200/csrc on the CPU.
//
// The CUDA kernels of the product are thin wrappers around the per-stream bodies in afsim_render.h.
// This harness compiles the same bodies (and the same host planner) with g++ and walks the same
// chunk x stage schedule with plain loops, so that chunking, state parking, ring addressing and the
// finalize reduction can be checked against the oracle in a container without a GPU.  It is never
// linked into libafsim.so and nothing in audio_forge_b200/ loads it.
//
// Build: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -shared -fPIC (tests/hostsim/Makefile)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../audio_forge_b200/csrc/afsim_plan.h"
#include "../../audio_forge_b200/csrc/afsim_eqscan.h"
#include "../../audio_forge_b200/csrc/afsim_render.h"

using namespace afsim;

namespace {
const float kFir[4][32] = {
#include "../../audio_forge_b200/csrc/true_peak_fir.inc"
};
thread_local std::string g_error;
}  // namespace

struct HostOptions {
    int block_samples = 0;               // 0: the rate's analysis block
    const double* const* vad = nullptr;  // per pair
    float* mk_rows_out = nullptr;        // [3][n_rows][n_pairs]
};

int run_hostsim(const std::vector<CandidatePlan>& plans, const float* const* passages, const size_t* passage_len,
                size_t n_passages, double fs, const uint32_t* pair_passage, const uint32_t* pair_candidate, size_t n_pairs,
                int chunk, int slots, int eq_k, int split, AfChainMetrics* out_metrics, float* out_audio, float* out_rows,
                const HostOptions& opt);

extern "C" {

const char* hostsim_last_error() { return g_error.c_str(); }

// All pairs must share structure / lookahead / length (one batch).  chunk / slots / eq_k select the
// schedule to exercise.  out_audio: nullptr or [n_pairs][T].
int hostsim_chain_sweep(const float* const* passages, const size_t* passage_len, size_t n_passages, double fs,
                        const AfCandidate* candidates, size_t n_candidates, const uint32_t* pair_passage,
                        const uint32_t* pair_candidate, size_t n_pairs, int chunk, int slots, int eq_k, int split,
                        AfChainMetrics* out_metrics, float* out_audio, float* out_rows /* [4][n_rows][n_pairs] or null */) {
    g_error.clear();
    std::vector<CandidatePlan> plans(n_candidates);
    for (size_t c = 0; c < n_candidates; ++c) {
        const int rc = plan_candidate(candidates[c].bands, candidates[c].settings, fs, &plans[c], &g_error);
        if (rc != AFSIM_OK) return rc;
    }
    return run_hostsim(plans, passages, passage_len, n_passages, fs, pair_passage, pair_candidate, n_pairs, chunk, slots, eq_k,
                       split, out_metrics, out_audio, out_rows, HostOptions());
}

// simulate_auto_makeup_control through the stage bodies: one capture, 480-sample control blocks.
// traces: [6][block_count] as afsim_auto_makeup_control.
int hostsim_makeup_control(const float* audio, size_t n, double fs, const double* vad, size_t n_vad, double noise_floor_db,
                           double noise_reliability, const AfAutoMakeupSettings* settings, int chunk, int slots, int direct,
                           float* traces, float* out_audio) {
    g_error.clear();
    std::vector<CandidatePlan> plans(1);
    const int rc = plan_makeup_control(*settings, fs, noise_floor_db, noise_reliability, n_vad != 0, &plans[0], &g_error);
    if (rc != AFSIM_OK) return rc;
    const size_t n_rows = (n + 479) / 480;
    if (n_rows == 0) return AFSIM_OK;
    const float* passages[1] = {audio};
    const size_t lens[1] = {n};
    const uint32_t zero[1] = {0};
    const double* vads[1] = {n_vad ? vad : nullptr};
    std::vector<float> rows(4 * n_rows), mk_rows(3 * n_rows);
    AfChainMetrics metrics;
    HostOptions opt;
    opt.block_samples = 480;
    opt.vad = vads;
    opt.mk_rows_out = mk_rows.data();
    const int rc2 = run_hostsim(plans, passages, lens, 1, fs, zero, zero, 1, chunk, slots, 5, direct ? 8 : 0, &metrics, out_audio,
                                rows.data(), opt);
    if (rc2 != AFSIM_OK) return rc2;
    for (size_t r = 0; r < n_rows; ++r) {
        traces[0 * n_rows + r] = mk_rows[0 * n_rows + r];
        traces[1 * n_rows + r] = mk_rows[1 * n_rows + r];
        traces[2 * n_rows + r] = mk_rows[2 * n_rows + r];
        traces[3 * n_rows + r] = rows[2 * n_rows + r];
        traces[4 * n_rows + r] = rows[0 * n_rows + r];
        traces[5 * n_rows + r] = rows[1 * n_rows + r];
    }
    return AFSIM_OK;
}

}  // extern "C"

int run_hostsim(const std::vector<CandidatePlan>& plans, const float* const* passages, const size_t* passage_len,
                size_t n_passages, double fs, const uint32_t* pair_passage, const uint32_t* pair_candidate, size_t n_pairs,
                int chunk, int slots, int eq_k, int split, AfChainMetrics* out_metrics, float* out_audio, float* out_rows,
                const HostOptions& opt) {
    const size_t n_candidates = plans.size();
    if (n_pairs == 0) return AFSIM_OK;
    const RateConstants rate = rate_constants(fs);
    const CandidatePlan& first = plans[pair_candidate[0]];
    const int T = static_cast<int>(passage_len[pair_passage[0]]);
    for (size_t i = 0; i < n_pairs; ++i) {
        const CandidatePlan& pl = plans[pair_candidate[i]];
        if (pl.structure != first.structure || pl.lookahead != first.lookahead || pl.input_stage != first.input_stage ||
            static_cast<int>(passage_len[pair_passage[i]]) != T) {
            g_error = "hostsim: pairs must form one batch";
            return AFSIM_INVALID_ARGUMENT;
        }
    }
    std::vector<uint64_t> passage_off(n_passages + 1, 0);
    for (size_t p = 0; p < n_passages; ++p) passage_off[p + 1] = passage_off[p] + passage_len[p];
    std::vector<float> signals(passage_off[n_passages] + 1);
    for (size_t p = 0; p < n_passages; ++p)
        std::memcpy(signals.data() + passage_off[p], passages[p], passage_len[p] * sizeof(float));
    std::vector<CandidateParams> params(n_candidates);
    for (size_t c = 0; c < n_candidates; ++c) params[c] = plans[c].params;

    const int S = static_cast<int>(n_pairs);
    const int S_pad = (S + 31) / 32 * 32;
    const size_t sp = static_cast<size_t>(S_pad);
    BatchArgs a{};
    a.structure = first.structure;
    a.lookahead = static_cast<int>(first.lookahead);
    a.input_stage = static_cast<int>(first.input_stage);
    a.n_streams = S;
    a.stride = S_pad;
    a.n_samples = T;
    a.block_samples = opt.block_samples > 0 ? opt.block_samples : rate.block_samples;
    a.fade_samples = rate.fade_samples;
    a.n_rows = (T + a.block_samples - 1) / a.block_samples;
    a.n_pad = 2;
    while (a.n_pad < a.n_rows) a.n_pad <<= 1;
    chunk = std::max(chunk, std::max(rate.fade_samples, a.lookahead + 1));
    chunk = (chunk + 7) / 8 * 8;
    if (a.input_stage == AF_INPUT_CLEANUP_GENTLE || a.input_stage == AF_INPUT_CLEANUP_STRONG)
        chunk = (chunk + kInputBlock - 1) / kInputBlock * kInputBlock;
    const bool auto_makeup = (a.structure & ST_AUTO_MAKEUP) != 0;
    MakeupConst mc{};
    if (auto_makeup) {  // whole blocks per chunk, as the product's build_sweep
        int unit = a.block_samples, eight = 8;
        while (eight) {
            const int t = unit % eight;
            unit = eight;
            eight = t;
        }
        unit = a.block_samples / unit * 8;
        chunk = (chunk + unit - 1) / unit * unit;
        mc = makeup_constants(fs, a.block_samples, T);
    }
    slots = std::max(slots, 2);
    a.ring_rows = slots * chunk;

    std::vector<uint32_t> cand(S_pad, 0), pair(S_pad, 0);
    std::vector<uint64_t> src_off(S_pad, 0), audio_off(S_pad, 0);
    uint32_t max_sections = 0;
    for (int s = 0; s < S; ++s) {
        cand[s] = pair_candidate[s];
        pair[s] = static_cast<uint32_t>(s);
        src_off[s] = passage_off[pair_passage[s]];
        audio_off[s] = static_cast<uint64_t>(s) * T;
        max_sections = std::max(max_sections, params[cand[s]].n_sections);
    }
    std::vector<float> buf_a(static_cast<size_t>(a.ring_rows) * sp), buf_b(static_cast<size_t>(a.ring_rows) * sp);
    std::vector<float> lim_sfx(static_cast<size_t>(a.lookahead + 1) * sp), rows(static_cast<size_t>(4) * std::max(a.n_rows, 1) * sp, 0.0f);
    std::vector<double> st_input(kStateInput * sp), st_de(kStateDeEsser * sp), st_eq(kStateEqPerSection * kMaxSections * sp),
        st_comp(kStateCompressor * sp), st_lim(kStateLimiter * sp), st_tp(kStateTruePeak * sp), de_tab(DE_FIELDS * sp);
    std::vector<double> w0(static_cast<size_t>(a.ring_rows) * sp), w1(w0.size()), w2(w0.size()), w3(w0.size()), w4(w0.size()),
        w5(w0.size()), w6(w0.size());
    std::vector<std::vector<double>> w_more(6, std::vector<double>(w0.size()));
    for (int i = 0; i < 6; ++i) a.w[7 + i] = w_more[i].data();
    std::vector<float> buf_c(static_cast<size_t>(a.ring_rows) * sp), buf_p(buf_c.size());
    a.w[0] = w0.data();
    a.w[1] = w1.data();
    a.w[2] = w2.data();
    a.w[3] = w3.data();
    a.w[4] = w4.data();
    a.w[5] = w5.data();
    a.w[6] = w6.data();
    a.stage_inputs = 1;
    a.buf_c = buf_c.data();
    a.buf_p = buf_p.data();
    std::vector<double> st_mk(kStateMakeup * sp), mk_ring(static_cast<size_t>(2 * std::max(mc.n_slots, 1) + 2 * kMaxMakeupSub) * sp, 0.0);
    std::vector<float> mk_rows(static_cast<size_t>(3) * std::max(a.n_rows, 1) * sp, 0.0f);
    std::vector<double> vad_pool;
    std::vector<int64_t> vad_off(S_pad, -1);
    if (auto_makeup) {
        a.st_mk = st_mk.data();
        a.mk_ring = mk_ring.data();
        a.mk_rows = mk_rows.data();
        a.mk_const = &mc;
        if (opt.vad) {
            for (int s = 0; s < S; ++s)
                if (opt.vad[s]) {
                    vad_off[s] = static_cast<int64_t>(vad_pool.size());
                    vad_pool.insert(vad_pool.end(), opt.vad[s], opt.vad[s] + a.n_rows);
                }
            vad_pool.push_back(0.0);
            a.mk_vad = vad_pool.data();
            a.mk_vad_off = vad_off.data();
        }
    }
    std::vector<StreamAccum> accum(sp);
    std::memset(accum.data(), 0, sp * sizeof(StreamAccum));
    std::vector<AfChainMetrics> metrics(n_pairs);
    a.params = params.data();
    a.cand = cand.data();
    a.pair = pair.data();
    a.src_off = src_off.data();
    a.audio_off = audio_off.data();
    a.signals = signals.data();
    a.audio = out_audio;
    a.buf_a = buf_a.data();
    a.buf_b = buf_b.data();
    a.lim_sfx = lim_sfx.data();
    a.st_input = st_input.data();
    a.st_deesser = st_de.data();
    a.st_eq = st_eq.data();
    a.st_comp = st_comp.data();
    a.st_lim = st_lim.data();
    a.st_tp = st_tp.data();
    a.rows = rows.data();
    a.accum = accum.data();
    a.eq_default = &rate.eq_default[0][0];
    a.cleanup = &rate.cleanup;
    a.de_tab = de_tab.data();
    std::vector<double> map_tab(static_cast<size_t>(MT_FIELDS) * sp);
    for (size_t s = 0; s < sp; ++s) fill_map_tab(map_tab.data(), sp, s, params[cand[s]]);
    a.map_tab = map_tab.data();
    a.metrics = metrics.data();

    // shared input stage (split bit 4): one render per distinct passage + fan-out, as build_sweep does for sweeps
    BatchArgs ua{};
    std::vector<uint32_t> uidx(S_pad, 0), ucand;
    std::vector<uint64_t> usrc;
    std::vector<float> ubuf, urows;
    std::vector<double> ustate;
    std::vector<StreamAccum> uaccum;
    const bool shared_eq = (split & 32) != 0;  // with bit 4: the EQ runs on the distinct (passage, EQ) pairs as well
    const bool shared_front = (split & 64) != 0 && shared_eq;  // ... and so does the compressor front (fused compressor)
    const bool shared_de = (split & 128) != 0;  // with bit 4: the de-esser's detector front runs on the distinct pairs
    std::vector<std::vector<double>> uw(7);
    std::vector<double> ust_comp, ust_de, ude_tab, umap_tab;
    std::vector<double> ust_eq;
    uint32_t shared_max_sections = 0;
    if (split & 16) {
        for (int s = 0; s < S; ++s) {
            size_t u = 0;
            while (u < usrc.size() &&
                   !(usrc[u] == src_off[s] &&
                     (!shared_de || (std::memcmp(params[ucand[u]].de + DE_DET, params[cand[s]].de + DE_DET, 30 * 8) == 0 &&
                                     params[ucand[u]].de[DE_DET_ATTACK] == params[cand[s]].de[DE_DET_ATTACK] &&
                                     params[ucand[u]].de[DE_DET_RELEASE] == params[cand[s]].de[DE_DET_RELEASE] &&
                                     std::memcmp(params[ucand[u]].de_det0, params[cand[s]].de_det0, sizeof params[0].de_det0) == 0)) &&
                     (!shared_eq || (std::memcmp(params[ucand[u]].eq, params[cand[s]].eq, sizeof params[0].eq) == 0 &&
                                     params[ucand[u]].n_sections == params[cand[s]].n_sections &&
                                     (params[ucand[u]].flags & (LF_EQ_FADE | LF_C_SIDECHAIN)) ==
                                         (params[cand[s]].flags & (LF_EQ_FADE | LF_C_SIDECHAIN))))))
                ++u;
            if (u == usrc.size()) {
                usrc.push_back(src_off[s]);
                ucand.push_back(cand[s]);
            }
            uidx[s] = static_cast<uint32_t>(u);
        }
        const int U = static_cast<int>(usrc.size()), U_pad = (U + 31) / 32 * 32;
        for (int u = 0; u < U; ++u) shared_max_sections = std::max(shared_max_sections, params[ucand[u]].n_sections);
        ust_eq.assign(static_cast<size_t>(kStateEqPerSection * kMaxSections) * U_pad, 0.0);
        usrc.resize(U_pad, 0);
        ucand.resize(U_pad, 0);
        ubuf.assign(static_cast<size_t>(a.ring_rows) * U_pad, 0.0f);
        urows.assign(static_cast<size_t>(std::max(a.n_rows, 1)) * U_pad, 0.0f);
        ustate.assign(static_cast<size_t>(kStateInput) * U_pad, 0.0);
        uaccum.resize(U_pad);
        std::memset(uaccum.data(), 0, U_pad * sizeof(StreamAccum));
        ua = a;
        ua.n_streams = U;
        ua.stride = U_pad;
        ua.cand = ucand.data();
        umap_tab.assign(static_cast<size_t>(MT_FIELDS) * U_pad, 0.0);
        for (int u = 0; u < U_pad; ++u) fill_map_tab(umap_tab.data(), U_pad, u, params[ucand[u]]);
        ua.map_tab = umap_tab.data();
        ua.src_off = usrc.data();
        ua.buf_a = ubuf.data();
        ua.rows = urows.data();
        ua.accum = uaccum.data();
        ua.st_input = ustate.data();
        ua.st_eq = ust_eq.data();
        if (shared_front) {
            for (int k = 0; k < 4; ++k) {
                uw[k].assign(static_cast<size_t>(a.ring_rows) * U_pad, 0.0);
                ua.w[k] = uw[k].data();
            }
            ust_comp.assign(static_cast<size_t>(kStateCompressor) * U_pad, 0.0);
            ua.st_comp = ust_comp.data();
            a.in_det = ua.w[0];
            a.in_wdb = ua.w[1];
            a.in_ipk = ua.w[2];
        }
        if (shared_de) {
            for (int k = 0; k < 7; ++k) {
                uw[k].assign(static_cast<size_t>(a.ring_rows) * U_pad, 0.0);
                ua.w[k] = uw[k].data();
                a.in_de[k] = ua.w[k];
            }
            ust_de.assign(static_cast<size_t>(kStateDeEsser) * U_pad, 0.0);
            ude_tab.assign(static_cast<size_t>(DE_FIELDS) * U_pad, 0.0);
            ua.st_deesser = ust_de.data();
            ua.de_tab = ude_tab.data();
            for (int u = 0; u < U; ++u) body_expand_deesser(ua, u);
        }
        a.in_unique = uidx.data();
        a.in_src = ubuf.data();
        a.in_rows = urows.data();
        a.in_accum = uaccum.data();
        a.in_stride = U_pad;
    }
    if (a.structure & ST_DEESSER)
        for (int s = 0; s < S; ++s) body_expand_deesser(a, s);
    std::vector<unsigned char> staging_bytes(std::max(kStagingBytesPerLane, std::max(kDeRc1StagingBytesPerLane, kDeRc3StagingBytesPerLane)) + 64);
    auto run_deesser = [&](const ChunkArgs& ck) {
        const Staging st{(split & 8) ? nullptr : staging_bytes.data(), 1, 0, 0};  // split bit 3: direct (unstaged) loads
        if ((split & 16) && shared_de) {
            for (int u = 0; u < ua.n_streams; ++u) body_de_ra(ua, ck, u, st);
            for (int g = (ck.len + kDeMapGroup - 1) / kDeMapGroup; g >= 0; --g)
                for (int u = 0; u < ua.n_streams; ++u) body_de_mb(ua, ck, u, g);
        } else {
            for (int s = 0; s < S; ++s) body_de_ra(a, ck, s, st);
            for (int g = (ck.len + kDeMapGroup - 1) / kDeMapGroup; g >= 0; --g)
                for (int s = 0; s < S; ++s) body_de_mb(a, ck, s, g);
        }
        if (split & 256) {  // R_c1 cut three ways (afsim_deesser.h)
            for (int s = 0; s < S; ++s) body_de_rc1a(a, ck, s, st);
            for (int g = (ck.len + kDeTargetGroup - 1) / kDeTargetGroup; g >= 0; --g)
                for (int s = 0; s < S; ++s) body_de_mc1b(a, ck, s, g);
            for (int s = 0; s < S; ++s) body_de_rc1c(a, ck, s, st);
        } else {
            for (int s = 0; s < S; ++s) body_de_rc(a, ck, s, st);
        }
        for (int g = (ck.len + kDeRebuildGroup - 1) / kDeRebuildGroup; g >= 0; --g)
            for (int s = 0; s < S; ++s) body_de_mc2(a, ck, s, g);
        for (int s = 0; s < S; ++s) body_de_rc3(a, ck, s, st);
    };
    const Staging stg{staging_bytes.data(), 1, 0, 0};
    const int n_chunks = T > 0 ? (T + chunk - 1) / chunk : 0;
    auto run_eq = [&](const ChunkArgs& ck) {
        for (uint32_t f = 0; f < max_sections; f += eq_k)
            for (int s = 0; s < S; ++s) {
                if (eq_k == 10)
                    body_eq<10>(a, ck, s, static_cast<int>(f));
                else
                    body_eq<5>(a, ck, s, static_cast<int>(f));
            }
    };
    for (int c = 0; c < n_chunks; ++c) {
        ChunkArgs ck;
        ck.n0 = c * chunk;
        ck.len = std::min(chunk, T - ck.n0);
        ck.row0 = (c % slots) * chunk;
        if (split & 16) {
            for (int u = 0; u < ua.n_streams; ++u) {
                if (input_uses_cleanup(ua))
                    body_input_cleanup(ua, ck, u);
                else
                    body_input(ua, ck, u);
            }
            if (shared_eq)
                for (uint32_t f = 0; f < shared_max_sections; f += eq_k)
                    for (int u = 0; u < ua.n_streams; ++u) {
                        if (eq_k == 10)
                            body_eq<10>(ua, ck, u, static_cast<int>(f));
                        else
                            body_eq<5>(ua, ck, u, static_cast<int>(f));
                    }
            if (shared_front) {
                const Staging ust{staging_bytes.data(), 1, 0, 0};
                for (int u = 0; u < ua.n_streams; ++u) body_comp_r1(ua, ck, u, ust);
                for (int g = (ck.len + kCompMapGroup - 1) / kCompMapGroup; g >= 0; --g)
                    for (int u = 0; u < ua.n_streams; ++u) body_comp_m2(ua, ck, u, g);
            }
            for (int g = (ck.len + kFanoutGroup - 1) / kFanoutGroup; g >= 0; --g)
                for (int s = 0; s < S; ++s) body_input_fanout(a, ck, s, g);
        } else {
            for (int s = 0; s < S; ++s) {
                if (input_uses_cleanup(a))
                    body_input_cleanup(a, ck, s);
                else
                    body_input(a, ck, s);
            }
        }
        if ((split & 16) && shared_eq) {
            if (a.structure & ST_DEESSER) run_deesser(ck);
        } else if (a.structure & ST_EQ_BEFORE_DEESSER) {
            run_eq(ck);
            if (a.structure & ST_DEESSER) run_deesser(ck);
        } else {
            if (a.structure & ST_DEESSER) run_deesser(ck);
            run_eq(ck);
        }
        const int n_groups = (ck.len + kGroup - 1) / kGroup + 1;  // one empty group past the end on purpose
        if (a.structure & ST_COMPRESSOR) {
            if ((split & 1) || auto_makeup) {
                const int n_cgroups = (ck.len + kCompMapGroup - 1) / kCompMapGroup + 1;
                if (!((split & 16) && shared_front)) {  // else: the shared front already ran on the distinct pairs
                    for (int s = 0; s < S; ++s) body_comp_r1(a, ck, s, stg);
                    for (int g = n_cgroups - 1; g >= 0; --g)  // any order: the maps are independent
                        for (int s = 0; s < S; ++s) body_comp_m2(a, ck, s, g);
                }
                for (int s = 0; s < S; ++s) body_comp_r3(a, ck, s, stg);
                for (int g = 0; g < n_cgroups; ++g)
                    for (int s = 0; s < S; ++s) body_comp_m4(a, ck, s, g);
                for (int s = 0; s < S; ++s) body_comp_r5(a, ck, s, stg);
                for (int g = n_cgroups - 1; g >= 0; --g)
                    for (int s = 0; s < S; ++s) body_comp_m6(a, ck, s, g);
                if (auto_makeup) {
                    const Staging st7{(split & 8) ? nullptr : staging_bytes.data(), 1, 0, 0};
                    for (int s = 0; s < S; ++s) body_comp_r7(a, ck, s, st7);
                }
            } else if ((split & 16) && shared_front) {
                for (int s = 0; s < S; ++s) body_compressor_shared(a, ck, s);
            } else {
                for (int s = 0; s < S; ++s) body_compressor(a, ck, s);
            }
        }
        if (a.structure & ST_LIMITER) {
            if (split & 2) {
                for (int g = (ck.len + kLimGroup - 1) / kLimGroup; g >= 0; --g)
                    for (int s = 0; s < S; ++s) body_lim_m(a, ck, s, g);
                for (int s = 0; s < S; ++s) body_lim_r(a, ck, s, stg);
            } else {
                for (int s = 0; s < S; ++s) body_limiter(a, ck, s);
            }
            if (split & 4) {
                for (int g = n_groups - 1; g >= 0; --g)
                    for (int s = 0; s < S; ++s) body_tp_fir_in(a, ck, s, g, kFir);
                for (int s = 0; s < S; ++s) body_tp_r(a, ck, s, stg);
                for (int g = n_groups - 1; g >= 0; --g)
                    for (int s = 0; s < S; ++s) body_tp_fir_out(a, ck, s, g, kFir);
            } else {
                for (int s = 0; s < S; ++s) body_output<true>(a, ck, s, kFir);
            }
        } else {
            for (int s = 0; s < S; ++s) body_output<false>(a, ck, s, kFir);
        }
    }
    std::vector<float> ws(finalize_workspace_floats(a.n_rows, a.n_pad));
    const Coop co{0, 1};
    for (int s = 0; s < S; ++s) body_finalize(a, s, co, ws.data());
    std::memcpy(out_metrics, metrics.data(), n_pairs * sizeof(AfChainMetrics));
    if (out_rows)
        for (int k = 0; k < 4; ++k)
            for (int r = 0; r < a.n_rows; ++r)
                for (int s = 0; s < S; ++s)
                    out_rows[(static_cast<size_t>(k) * a.n_rows + r) * S + s] = rows[(static_cast<size_t>(k) * a.n_rows + r) * sp + s];
    if (opt.mk_rows_out)
        for (int k = 0; k < 3; ++k)
            for (int r = 0; r < a.n_rows; ++r)
                for (int s = 0; s < S; ++s)
                    opt.mk_rows_out[(static_cast<size_t>(k) * a.n_rows + r) * S + s] =
                        mk_rows[(static_cast<size_t>(k) * a.n_rows + r) * sp + s];
    return AFSIM_OK;
}

extern "C" {

// Piece index of every stream of a batch (afsim_plan.cpp, cut_stream_group); returns the number of pieces.
int hostsim_cut_group(const uint32_t* passage, const uint32_t* eq_class, size_t n, int max_streams, uint32_t* out_piece,
                      uint32_t* out_position) {
    const std::vector<uint32_t> p(passage, passage + n), c(eq_class, eq_class + n);
    const std::vector<std::vector<uint32_t>> pieces = cut_stream_group(p, c, max_streams);
    for (size_t k = 0; k < pieces.size(); ++k)
        for (size_t j = 0; j < pieces[k].size(); ++j) {
            out_piece[pieces[k][j]] = static_cast<uint32_t>(k);
            out_position[pieces[k][j]] = static_cast<uint32_t>(j);
        }
    return static_cast<int>(pieces.size());
}

// Planner outputs, for coefficient-level tests against the oracle.
int hostsim_plan(const AfBand* bands, const AfChainSettings* settings, double fs, CandidateParams* out, uint32_t* structure,
                 uint32_t* lookahead) {
    CandidatePlan plan;
    const int rc = plan_candidate(bands, *settings, fs, &plan, &g_error);
    if (rc != AFSIM_OK) return rc;
    *out = plan.params;
    *structure = plan.structure;
    *lookahead = plan.lookahead;
    return AFSIM_OK;
}
size_t hostsim_candidate_params_size() { return sizeof(CandidateParams); }

// The time-parallel EQ render of one passage (afsim_eqscan.h) walked on the host with the kernel's structure:
// transposed segments, zero-state local pass, the 1024-thread block scan (per-thread composites, Kogge-Stone inside
// each 32-lane group, group totals scanned the same way), apply pass fused with the next section's local pass.
int hostsim_eq_scan(const float* audio, size_t n, double fs, const AfBand* bands, int log2_len, float* out) {
    g_error.clear();
    CandidatePlan plan;
    const int rc = plan_eq_only(bands, fs, &plan, &g_error);
    if (rc != AFSIM_OK) return rc;
    if (n == 0) return AFSIM_OK;
    const int n_sections = static_cast<int>(plan.params.n_sections);
    if (n_sections == 0) {
        std::memcpy(out, audio, n * sizeof(float));
        return AFSIM_OK;
    }
    const size_t len = size_t(1) << log2_len, n_seg = (n + len - 1) / len;
    std::vector<float> xt(n_seg * len, 0.0f);
    for (size_t i = 0; i < n; ++i) xt[(i % len) * n_seg + i / len] = audio[i];
    std::vector<double> end(2 * n_seg), init(2 * n_seg);
    const Bq first = bq_from(plan.params.eq[0]);
    for (size_t k = 0; k < n_seg; ++k)
        eqscan_segment<false, true>(xt.data(), n_seg, k, static_cast<int>(len), first, 0.0, 0.0, first, &end[k], &end[n_seg + k]);
    constexpr int T = 1024;
    for (int j = 0; j < n_sections; ++j) {
        const Bq cur = bq_from(plan.params.eq[j]);
        double m[4];
        biquad_transition_power(cur, log2_len, m);
        const size_t per = (n_seg + T - 1) / T;
        std::vector<Affine2> incl(T), tot(T / 32);
        for (int t = 0; t < T; ++t) {
            Affine2 comp = affine_identity();
            for (size_t k = t * per; k < std::min(n_seg, (t + 1) * per); ++k)
                comp = affine_then(comp, Affine2{m[0], m[1], m[2], m[3], end[k], end[n_seg + k]});
            incl[t] = comp;
        }
        auto group_scan = [](Affine2* v) {  // Kogge-Stone over 32 lanes
            for (int d = 1; d < 32; d <<= 1) {
                Affine2 prev[32];
                for (int l = 0; l < 32; ++l) prev[l] = l >= d ? v[l - d] : v[l];
                for (int l = d; l < 32; ++l) v[l] = affine_then(prev[l], v[l]);
            }
        };
        for (int w = 0; w < T / 32; ++w) {
            group_scan(&incl[w * 32]);
            tot[w] = incl[w * 32 + 31];
        }
        group_scan(tot.data());
        for (int t = 0; t < T; ++t) {
            const int lane = t & 31, warp = t >> 5;
            Affine2 excl = lane == 0 ? affine_identity() : incl[t - 1];
            if (warp > 0) excl = affine_then(tot[warp - 1], excl);
            double s1 = excl.v0, s2 = excl.v1;
            for (size_t k = t * per; k < std::min(n_seg, (t + 1) * per); ++k) {
                init[k] = s1;
                init[n_seg + k] = s2;
                const double t1 = m[0] * s1 + m[1] * s2 + end[k];
                const double t2 = m[2] * s1 + m[3] * s2 + end[n_seg + k];
                s1 = t1;
                s2 = t2;
            }
        }
        const bool last = j + 1 == n_sections;
        const Bq nxt = last ? cur : bq_from(plan.params.eq[j + 1]);
        for (size_t k = 0; k < n_seg; ++k) {
            if (last)
                eqscan_segment<true, false>(xt.data(), n_seg, k, static_cast<int>(len), cur, init[k], init[n_seg + k], nxt, nullptr, nullptr);
            else
                eqscan_segment<true, true>(xt.data(), n_seg, k, static_cast<int>(len), cur, init[k], init[n_seg + k], nxt, &end[k], &end[n_seg + k]);
        }
    }
    for (size_t i = 0; i < n; ++i) out[i] = xt[(i % len) * n_seg + i / len];
    return AFSIM_OK;
}

}  // extern "C"
