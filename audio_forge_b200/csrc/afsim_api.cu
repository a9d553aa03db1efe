// afsim_api.cu -- the C ABI of include/afsim.h: handle, batching, the chunk x stage wavefront.
//
// Host orchestration only; all signal arithmetic is in the kernels (afsim_kernels.cu) and all
// settings arithmetic in the planner (afsim_plan.cpp).  There is no CPU render path in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <tuple>
#include <type_traits>
#include <vector>

#include "../../include/afsim.h"
#include "afsim_kernels.h"
#include "afsim_plan.h"
#include "afsim_render.h"
#include "afsim_resample.h"
#include "afsim_tail.h"

using namespace afsim;

namespace {
thread_local std::string g_create_error;
constexpr int kMaxStages = 32;
}  // namespace

// Device allocations are recycled through the handle: a sweep's work rings are GBs, and cudaMalloc /
// cudaFree of those on every call would dominate the host-buffer (end-to-end) path.
struct DevicePool {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;
    size_t cached_bytes = 0;
    size_t max_cached = size_t(96) << 30;  // AFSIM_POOL_MAX_GB at afsim_create; afsim_trim() gives everything back
    // best fit: the smallest cached block that holds `bytes` and wastes at most a quarter of itself
    cudaError_t take(size_t bytes, void** out, size_t* actual) {
        *actual = bytes;
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = free_blocks.lower_bound(bytes);
            if (it != free_blocks.end() && it->first - bytes <= it->first / 4) {
                *out = it->second;
                *actual = it->first;
                cached_bytes -= it->first;
                free_blocks.erase(it);
                return cudaSuccess;
            }
        }
        cudaError_t err = cudaMalloc(out, bytes);
        if (err == cudaErrorMemoryAllocation) {  // give the cache back and retry once
            cudaGetLastError();
            trim();
            err = cudaMalloc(out, bytes);
        }
        return err;
    }
    void give(size_t bytes, void* p) {
        std::lock_guard<std::mutex> lock(mu);
        if (cached_bytes + bytes > max_cached) {
            cudaFree(p);
            return;
        }
        free_blocks.emplace(bytes, p);
        cached_bytes += bytes;
    }
    void trim() {
        std::lock_guard<std::mutex> lock(mu);
        for (auto& kv : free_blocks) cudaFree(kv.second);
        free_blocks.clear();
        cached_bytes = 0;
    }
};

struct AfsimSweep;

// The last resampler plan of a handle stays on the device: the tool's cases repeat one (configuration, length) pair
// for many signals, and a 60 s plan is 2.9 M frames (46 MB) walked sequentially on the host.
struct ResampleCache {
    ResamplePlan plan;
    ResampleFrame* d_frames = nullptr;
    double* d_table = nullptr;
    // host-buffer path: two pinned staging areas (inputs + outputs of one group of signals each) and their "drained" events
    void* pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes = 0;
    cudaEvent_t done[2] = {nullptr, nullptr};
    bool matches(const AfResamplerSpec& sp, size_t n_in) const {
        return d_frames && plan.n_in == n_in && std::memcmp(&plan.spec, &sp, sizeof sp) == 0;
    }
    void drop() {
        if (d_frames) cudaFree(d_frames);
        if (d_table) cudaFree(d_table);
        d_frames = nullptr;
        d_table = nullptr;
    }
    void drop_staging() {
        for (int i = 0; i < 2; ++i) {
            if (pinned[i]) cudaFreeHost(pinned[i]);
            if (done[i]) cudaEventDestroy(done[i]);
            pinned[i] = nullptr;
            done[i] = nullptr;
        }
        pinned_bytes = 0;
    }
};
struct AfsimHandle {
    int device = 0;
    DevicePool pool;
    cudaStream_t stream = nullptr;   // caller-visible stream: every call starts and ends on it
    bool own_stream = false;
    cudaStream_t stage_stream[kMaxStages] = {};      // high priority: serial (latency-critical) stage kernels
    cudaStream_t stage_stream_map[kMaxStages] = {};  // low priority: map kernels that fill every SM
    cudaEvent_t ev_fork = nullptr;
    std::string error;
    std::recursive_mutex call_mu;     // one call at a time per handle (include/afsim.h: threading contract)
    std::set<AfsimSweep*> live;       // sweeps whose buffers come from `pool`: detached by afsim_destroy
    ResampleCache resample;
};

namespace {

struct DeviceBuffers {  // returns what it allocated to the handle's pool
    DevicePool* pool = nullptr;
    std::vector<std::pair<size_t, void*>> blocks;
    ~DeviceBuffers() {
        for (auto& b : blocks) {
            if (pool)
                pool->give(b.first, b.second);
            else
                cudaFree(b.second);
        }
    }
    template <typename T>
    cudaError_t alloc(T** out, size_t count) {
        void* p = nullptr;
        size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) / 256 * 256;
        const cudaError_t err = pool ? pool->take(bytes, &p, &bytes) : cudaMalloc(&p, bytes);
        if (err != cudaSuccess) return err;
        blocks.emplace_back(bytes, p);
        *out = static_cast<T*>(p);
        return cudaSuccess;
    }
};

enum StageKind { SK_INPUT, SK_INPUT_TP, SK_DEESSER, SK_EQ, SK_COMPRESSOR, SK_LIMITER, SK_OUTPUT, SK_SPLIT, SK_INPUT_SHARED, SK_INPUT_FANOUT, SK_EQ_SHARED, SK_SPLIT_SHARED, SK_TAIL };
struct StageDesc {
    StageKind kind;
    int arg;  // SK_EQ: first section; SK_SPLIT: SplitOp
};

struct Batch {
    BatchArgs args{};
    BatchArgs shared_input{};       // the distinct passages' input stage (n_streams == 0: every stream renders its own)
    bool shared_deesser = false;    // ... followed by the de-esser's detector front
    size_t shared_rows_elems = 0;
    std::vector<StageDesc> stages;
    std::vector<uint32_t> members;  // caller's pair index of stream s
    size_t mk_ring_elems = 0;       // auto makeup: doubles in args.mk_ring
    int chunk = 0, slots = 0, eq_k = 0;
    TailMap tail_map{};             // SK_TAIL: tensor map of the tail's input ring (buf_a)
    int* tail_err = nullptr;        // SK_TAIL: the sweep's watchdog word
    std::vector<cudaEvent_t> events;  // [stage][slot]
    ~Batch() {
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
};

}  // namespace

struct AfsimSweep {
    DeviceBuffers mem;
    std::vector<std::unique_ptr<Batch>> batches;
    AfChainMetrics* d_metrics = nullptr;   // [n_pairs], caller's pair order
    int* d_tail_err = nullptr;             // set by the fused tail kernel if its pipeline watchdog fired
    StreamAccum* d_accum_first = nullptr;  // accum table of batch 0 (afsim_eq_render reads stream 0)
    float* d_audio = nullptr;
    std::vector<uint64_t> audio_off;       // per pair
    std::vector<uint64_t> pair_len;
    size_t n_pairs = 0;
    int kernels_per_launch = 0;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    bool launched = false;
    AfsimHandle* owner = nullptr;  // nullptr once the handle is gone (afsim_destroy detaches live sweeps)
    int device = 0;
    bool quiesced = false;         // afsim_sweep_release synchronised the streams this sweep ran on
    ~AfsimSweep() {
        // error paths (a failed build / launch) get here with copies or kernels possibly still queued: nothing may
        // hand these buffers to the next sweep before the device is idle
        if (!quiesced) {
            cudaSetDevice(device);
            cudaDeviceSynchronize();
        }
        if (ev_start) cudaEventDestroy(ev_start);
        if (ev_stop) cudaEventDestroy(ev_stop);
    }
};

namespace {

int set_error(AfsimHandle* h, int status, const std::string& msg) {
    if (h) h->error = msg;
    return status;
}
int cuda_fail(AfsimHandle* h, cudaError_t err, const char* what) {
    const int status = err == cudaErrorMemoryAllocation ? AFSIM_OUT_OF_MEMORY : AFSIM_CUDA_ERROR;
    if (err == cudaErrorMemoryAllocation) cudaGetLastError();  // clear the sticky-free error
    return set_error(h, status, std::string(what) + ": " + cudaGetErrorString(err));
}
#define AF_CUDA(h, expr)                                         \
    do {                                                         \
        const cudaError_t af_err__ = (expr);                     \
        if (af_err__ != cudaSuccess) return cuda_fail((h), af_err__, #expr); \
    } while (0)

#define AF_LOCK(h) std::lock_guard<std::recursive_mutex> af_lock__((h)->call_mu)

int env_int(const char* name, int fallback) {
    const char* v = std::getenv(name);
    if (!v || !*v) return fallback;
    const long parsed = std::strtol(v, nullptr, 10);
    return parsed > 0 ? static_cast<int>(parsed) : fallback;
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ---- sweep construction ----------------------------------------------------------------------------------

struct PassageSource {
    const float* const* host = nullptr;  // host passages (nullptr: synthetic on device)
    int synth_kind = 0;
};

struct SweepOptions {
    int block_samples = 0;               // 0: the rate's analysis block (python_api.rs:512-513); 480 for the makeup control sim
    const double* const* vad = nullptr;  // per pair: per-block VAD probabilities or nullptr (simulate_auto_makeup_control)
    bool makeup_rows = false;            // keep the per-block makeup / activity / reliability traces
};

int lcm_int(int a, int b) {
    int x = a, y = b;
    while (y) {
        const int t = x % y;
        x = y;
        y = t;
    }
    return a / x * b;
}

int build_sweep(AfsimHandle* h, const PassageSource& src, const size_t* passage_len, size_t n_passages, double fs,
                const AfCandidate* candidates, size_t n_candidates, const CandidatePlan* preplanned,
                const uint32_t* pair_passage, const uint32_t* pair_candidate, size_t n_pairs, int want_audio,
                AfsimSweep** out_sweep, const SweepOptions& opt = SweepOptions()) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    if (!out_sweep) return set_error(h, AFSIM_INVALID_ARGUMENT, "out_sweep is null");
    *out_sweep = nullptr;
    if ((pair_passage == nullptr) != (pair_candidate == nullptr))
        return set_error(h, AFSIM_INVALID_ARGUMENT, "pair_passage and pair_candidate must both be given or both be null");
    if (n_pairs > 0 && (n_passages == 0 || n_candidates == 0))
        return set_error(h, AFSIM_INVALID_ARGUMENT, "a sweep needs at least one passage and one candidate");
    if (!pair_passage && n_pairs != n_passages * n_candidates)
        return set_error(h, AFSIM_INVALID_ARGUMENT, "n_pairs must equal n_candidates * n_passages for a full cross product");
    AF_CUDA(h, cudaSetDevice(h->device));

    // plan every candidate (reference constructor + setters -> constants)
    std::vector<CandidatePlan> plans(n_candidates);
    for (size_t c = 0; c < n_candidates; ++c) {
        if (preplanned) {
            plans[c] = preplanned[c];
        } else {
            std::string msg;
            const int rc = plan_candidate(candidates[c].bands, candidates[c].settings, fs, &plans[c], &msg);
            if (rc != AFSIM_OK) return set_error(h, rc, msg);
            if (plans[c].deesser_unstable)
                return set_error(h, AFSIM_UNSUPPORTED,
                                 "de-esser band edges reach Nyquist at this sample rate: the reference's de-esser is unstable there");
        }
    }
    const RateConstants rate = rate_constants(fs);

    auto sweep = std::make_unique<AfsimSweep>();
    sweep->mem.pool = &h->pool;
    sweep->device = h->device;
    sweep->n_pairs = n_pairs;

    // passages -> one device pool
    std::vector<uint64_t> passage_off(n_passages + 1, 0);
    for (size_t p = 0; p < n_passages; ++p) {
        if (passage_len[p] > 0x7fffff00u) return set_error(h, AFSIM_INVALID_ARGUMENT, "passage too long");
        passage_off[p + 1] = passage_off[p] + passage_len[p];
    }
    float* d_signals = nullptr;
    AF_CUDA(h, sweep->mem.alloc(&d_signals, passage_off[n_passages]));
    if (src.host) {
        for (size_t p = 0; p < n_passages; ++p)
            if (passage_len[p])
                AF_CUDA(h, cudaMemcpyAsync(d_signals + passage_off[p], src.host[p], passage_len[p] * sizeof(float),
                                           cudaMemcpyHostToDevice, h->stream));
    } else if (n_passages) {
        for (size_t p = 1; p < n_passages; ++p)
            if (passage_len[p] != passage_len[0])
                return set_error(h, AFSIM_INVALID_ARGUMENT, "synthetic passages must share one length");
        AF_CUDA(h, launch_synth(d_signals, passage_len[0], static_cast<int>(n_passages), src.synth_kind, fs, h->stream));
    }

    // candidate constants
    CandidateParams* d_params = nullptr;
    AF_CUDA(h, sweep->mem.alloc(&d_params, n_candidates));
    {
        std::vector<CandidateParams> host_params(n_candidates);
        for (size_t c = 0; c < n_candidates; ++c) host_params[c] = plans[c].params;
        if (n_candidates)
            AF_CUDA(h, cudaMemcpyAsync(d_params, host_params.data(), n_candidates * sizeof(CandidateParams),
                                       cudaMemcpyHostToDevice, h->stream));
        AF_CUDA(h, cudaStreamSynchronize(h->stream));  // host_params goes out of scope
    }
    double* d_eq_default = nullptr;
    AF_CUDA(h, sweep->mem.alloc(&d_eq_default, 50));
    AF_CUDA(h, cudaMemcpyAsync(d_eq_default, rate.eq_default, sizeof rate.eq_default, cudaMemcpyHostToDevice, h->stream));
    CleanupConst* d_cleanup = nullptr;
    AF_CUDA(h, sweep->mem.alloc(&d_cleanup, 1));
    AF_CUDA(h, cudaMemcpyAsync(d_cleanup, &rate.cleanup, sizeof rate.cleanup, cudaMemcpyHostToDevice, h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));

    AF_CUDA(h, sweep->mem.alloc(&sweep->d_tail_err, 1));
    AF_CUDA(h, cudaMemsetAsync(sweep->d_tail_err, 0, sizeof(int), h->stream));
    AF_CUDA(h, sweep->mem.alloc(&sweep->d_metrics, n_pairs));
    AF_CUDA(h, cudaMemsetAsync(sweep->d_metrics, 0, std::max<size_t>(n_pairs, 1) * sizeof(AfChainMetrics), h->stream));

    // output audio pool, caller's pair order
    sweep->audio_off.assign(n_pairs + 1, 0);
    sweep->pair_len.assign(n_pairs, 0);
    auto passage_of = [&](size_t i) { return pair_passage ? pair_passage[i] : static_cast<uint32_t>(i % n_passages); };
    auto candidate_of = [&](size_t i) { return pair_candidate ? pair_candidate[i] : static_cast<uint32_t>(i / n_passages); };
    for (size_t i = 0; i < n_pairs; ++i) {
        if (passage_of(i) >= n_passages || candidate_of(i) >= n_candidates)
            return set_error(h, AFSIM_INVALID_ARGUMENT, "pair index out of range");
        sweep->pair_len[i] = passage_len[passage_of(i)];
        sweep->audio_off[i + 1] = sweep->audio_off[i] + (want_audio ? sweep->pair_len[i] : 0);
    }
    if (want_audio) AF_CUDA(h, sweep->mem.alloc(&sweep->d_audio, sweep->audio_off[n_pairs]));

    // batches: streams that share structure, lookahead, input stage and length
    typedef std::tuple<uint32_t, uint32_t, uint32_t, uint64_t> Key;
    std::map<Key, std::vector<uint32_t>> groups;
    for (size_t i = 0; i < n_pairs; ++i) {
        const CandidatePlan& pl = plans[candidate_of(i)];
        groups[Key(pl.structure, pl.lookahead, pl.input_stage, sweep->pair_len[i])].push_back(static_cast<uint32_t>(i));
    }
    // EQ class of a candidate: same sections, coefficients, fade flag (and sidechain flag, for the compressor front)
    auto eq_class_key = [](const CandidateParams& p) {
        std::vector<unsigned char> key(sizeof p.eq + 8);
        std::memcpy(key.data(), p.eq, sizeof p.eq);
        const uint32_t meta[2] = {p.n_sections, p.flags & (LF_EQ_FADE | LF_C_SIDECHAIN)};
        std::memcpy(key.data() + sizeof p.eq, meta, 8);
        return key;
    };
    // Large compressor grids over few (passage, EQ) pairs are cut into pieces of up to AFSIM_SUBBATCH streams: a piece
    // runs the R/M split kernels with the shared EQ + compressor front, which is faster than the fused kernels on the
    // whole group (C3, 131072 streams: 19.9 -> see DESIGN section 6).  The pieces run one after another (run_batch
    // joins on the handle's stream), so they share one set of ring buffers (`ring_scratch`).
    struct Piece {
        Key key;
        std::vector<uint32_t> members;
        bool share_rings = false, first = false;
    };
    std::vector<Piece> pieces;
    {
        const int sub = env_int("AFSIM_SUBBATCH", 16384);
        const int split_mode = env_int("AFSIM_SPLIT", 0);
        for (auto& kv : groups) {
            std::vector<uint32_t>& members = kv.second;
            const uint32_t structure = std::get<0>(kv.first);
            const int S = static_cast<int>(members.size());
            bool cut = S > sub && sub <= 16384 && split_mode != 1 && env_int("AFSIM_SHARED_INPUT", 1) == 1 &&
                       (structure & ST_COMPRESSOR) && (structure & ST_EQ) &&
                       !(structure & (ST_DEESSER | ST_AUTO_MAKEUP | ST_INPUT_TRUE_PEAK));
            if (!cut) {
                pieces.push_back({kv.first, std::move(members), false, false});
                continue;
            }
            // EQ class of every stream; few distinct (passage, EQ) pairs -> pieces (afsim_plan.cpp)
            std::map<std::vector<unsigned char>, uint32_t> classes;
            std::vector<int64_t> class_of(plans.size(), -1);
            std::vector<uint32_t> m_passage(S), m_class(S);
            for (int k = 0; k < S; ++k) {
                const uint32_t c = candidate_of(members[k]);
                if (class_of[c] < 0)
                    class_of[c] = classes.emplace(eq_class_key(plans[c].params), static_cast<uint32_t>(classes.size())).first->second;
                m_passage[k] = passage_of(members[k]);
                m_class[k] = static_cast<uint32_t>(class_of[c]);
            }
            const std::vector<std::vector<uint32_t>> cuts = cut_stream_group(m_passage, m_class, sub);
            if (cuts.size() == 1) {
                pieces.push_back({kv.first, std::move(members), false, false});
                continue;
            }
            for (size_t k = 0; k < cuts.size(); ++k) {
                Piece pc;
                pc.key = kv.first;
                pc.members.reserve(cuts[k].size());
                for (uint32_t pos : cuts[k]) pc.members.push_back(members[pos]);
                pc.share_rings = true;
                pc.first = k == 0;
                pieces.push_back(std::move(pc));
            }
        }
    }
    std::vector<std::pair<size_t, void*>> ring_scratch;  // ring buffers shared by the pieces of one cut group
    size_t ring_next = 0;
    for (Piece& piece : pieces) {
        // Streams with the same EQ section count share warps: an EQ slice past a warp's last section is skipped
        // by the whole warp (body_eq returns early), instead of running as a predicated pass-through.
        std::vector<uint32_t>& members = piece.members;
        const Key& piece_key = piece.key;
        if (piece.first) ring_scratch.clear();
        ring_next = 0;
        const bool share_rings = piece.share_rings;
        auto alloc_ring = [&](auto** out, size_t count) -> cudaError_t {
            typedef typename std::remove_pointer<typename std::remove_pointer<decltype(out)>::type>::type T;
            if (!share_rings) return sweep->mem.alloc(out, count);
            const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
            if (ring_next < ring_scratch.size() && ring_scratch[ring_next].first >= bytes) {
                *out = static_cast<T*>(ring_scratch[ring_next++].second);
                return cudaSuccess;
            }
            const cudaError_t err = sweep->mem.alloc(out, count);
            if (err != cudaSuccess) return err;
            if (ring_next < ring_scratch.size())
                ring_scratch[ring_next] = std::make_pair(bytes, static_cast<void*>(*out));
            else
                ring_scratch.emplace_back(bytes, static_cast<void*>(*out));
            ++ring_next;
            return cudaSuccess;
        };
        std::stable_sort(members.begin(), members.end(), [&](uint32_t x, uint32_t y) {
            return plans[candidate_of(x)].params.n_sections < plans[candidate_of(y)].params.n_sections;
        });
        auto batch = std::make_unique<Batch>();
        BatchArgs& a = batch->args;
        const int S = static_cast<int>(members.size());
        const int S_pad = round_up(S, 32);
        const int T = static_cast<int>(std::get<3>(piece_key));
        a.structure = std::get<0>(piece_key);
        a.lookahead = static_cast<int>(std::get<1>(piece_key));
        a.input_stage = static_cast<int>(std::get<2>(piece_key));
        a.n_streams = S;
        a.stride = S_pad;
        a.n_samples = T;
        a.block_samples = opt.block_samples > 0 ? opt.block_samples : rate.block_samples;
        a.fade_samples = rate.fade_samples;
        a.n_rows = (T + a.block_samples - 1) / a.block_samples;
        a.n_pad = 2;
        while (a.n_pad < a.n_rows) a.n_pad <<= 1;
        a.params = d_params;
        a.signals = d_signals;
        a.audio = sweep->d_audio;
        a.eq_default = d_eq_default;
        a.cleanup = d_cleanup;
        a.metrics = sweep->d_metrics;
        {  // tuning knobs are read HERE, once per sweep (never cached across sweeps: tests flip them between calls)
            const char* v = std::getenv("AFSIM_MAP_BLOCKS_PER_SM");
            a.map_blocks_per_sm = v && *v ? std::atoi(v) : 0;
            v = std::getenv("AFSIM_FIR_BLOCKS_PER_SM");
            a.fir_blocks_per_sm = v && *v ? std::atoi(v) : -1;
        }

        // chunking: a multiple of 8 (true-peak FIR groups) and of the compressor micro-tile, long
        // enough to hold the biquad crossfade and the limiter lookback
        int chunk = env_int("AFSIM_CHUNK", 1024);
        chunk = std::max(chunk, std::max(rate.fade_samples, a.lookahead + 1));
        chunk = round_up(chunk, 8);
        if (a.input_stage == AF_INPUT_CLEANUP_GENTLE || a.input_stage == AF_INPUT_CLEANUP_STRONG)
            chunk = round_up(chunk, kInputBlock);  // the cleanup stage works in whole 480-sample blocks
        const bool auto_makeup = (a.structure & ST_AUTO_MAKEUP) != 0;
        if (auto_makeup) {
            // the makeup steps at block ends and a block is fed to the loudness meter (or not) as a whole: chunks are
            // whole blocks (and still multiples of 8, and of 480 with the cleanup stage)
            int unit = lcm_int(a.block_samples, 8);
            if (a.input_stage == AF_INPUT_CLEANUP_GENTLE || a.input_stage == AF_INPUT_CLEANUP_STRONG) unit = lcm_int(unit, kInputBlock);
            if (unit > 16384)
                return set_error(h, AFSIM_UNSUPPORTED, "auto makeup with the input cleanup stage is not available at this sample rate");
            chunk = round_up(chunk, unit);
        }
        batch->chunk = chunk;
        const int n_chunks = T > 0 ? (T + chunk - 1) / chunk : 0;
        // ring slots = chunks in flight + 1; few-stream batches need the stage wavefront to fill the GPU
        // (a split batch runs up to ~24 stage kernels per chunk: one ring slot per stage keeps all of them busy)
        // (fused kernels, 16384 < S <= 32768: eight slots keep ~7 of the ~12 stage kernels of a chunk in flight -- C5 at 32768
        // streams per GPU 1840 -> 1679 ms; at 65536 streams the GPU is full with three and 8 slots are 90 GB of rings)
        int slots = S <= 16384 ? 26 : (S <= 32768 ? 8 : (S <= 65536 ? 4 : 2));
        slots = env_int("AFSIM_SLOTS", slots);
        slots = std::max(2, std::min(slots, std::max(2, n_chunks + 1)));
        batch->slots = slots;
        a.ring_rows = slots * chunk;

        // per-stream tables
        std::vector<uint32_t> cand(S_pad, 0), pair(S_pad, 0);
        std::vector<uint64_t> src_off(S_pad, 0), audio_off(S_pad, 0);
        uint32_t max_sections = 0;
        batch->members = members;
        for (int s = 0; s < S; ++s) {
            const uint32_t i = members[s];
            cand[s] = candidate_of(i);
            pair[s] = i;
            src_off[s] = passage_off[passage_of(i)];
            audio_off[s] = sweep->audio_off[i];
            max_sections = std::max(max_sections, plans[cand[s]].params.n_sections);
        }
        uint32_t *d_cand = nullptr, *d_pair = nullptr;
        uint64_t *d_src = nullptr, *d_aoff = nullptr;
        AF_CUDA(h, sweep->mem.alloc(&d_cand, S_pad));
        AF_CUDA(h, sweep->mem.alloc(&d_pair, S_pad));
        AF_CUDA(h, sweep->mem.alloc(&d_src, S_pad));
        AF_CUDA(h, sweep->mem.alloc(&d_aoff, S_pad));
        AF_CUDA(h, cudaMemcpyAsync(d_cand, cand.data(), S_pad * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        AF_CUDA(h, cudaMemcpyAsync(d_pair, pair.data(), S_pad * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        AF_CUDA(h, cudaMemcpyAsync(d_src, src_off.data(), S_pad * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
        AF_CUDA(h, cudaMemcpyAsync(d_aoff, audio_off.data(), S_pad * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
        AF_CUDA(h, cudaStreamSynchronize(h->stream));
        {  // the map kernels' constants, stream-minor (MapField)
            std::vector<double> tab(static_cast<size_t>(MT_FIELDS) * S_pad);
            for (int s = 0; s < S_pad; ++s) fill_map_tab(tab.data(), S_pad, s, plans[cand[s]].params);
            double* d_tab = nullptr;
            AF_CUDA(h, sweep->mem.alloc(&d_tab, tab.size()));
            AF_CUDA(h, cudaMemcpyAsync(d_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            AF_CUDA(h, cudaStreamSynchronize(h->stream));
            a.map_tab = d_tab;
        }
        a.cand = d_cand;
        a.pair = d_pair;
        a.src_off = d_src;
        a.audio_off = d_aoff;

        const size_t sp = static_cast<size_t>(S_pad);
        // Few-stream batches cannot fill the GPU with one thread per stream: cut the stages into serial
        // recurrences + parallel maps (afsim_split.h).  AFSIM_SPLIT: 1 = never, 2 = always, else by size.
        const int split_mode = env_int("AFSIM_SPLIT", 0);
        const bool split = split_mode == 2 || (split_mode != 1 && S <= 16384);
        const size_t ring_elems = static_cast<size_t>(a.ring_rows) * sp;
        AF_CUDA(h, alloc_ring(&a.buf_a, ring_elems));
        // The limiter / true-peak tail as ONE SM-local, TMA-fed kernel (afsim_tail.cu) instead of five stage kernels.
        // AFSIM_TAIL: 1 = never, 2 = always; default: the few-stream (split) batches whose chain IS the tail -- no
        // compressor, no de-esser (batch true-peak detection + limiting, BASELINE config 4).  Behind a compressor /
        // de-esser wavefront the five thin stage kernels overlap with the other stages' kernels across chunks, and one
        // CTA-per-32-streams kernel that holds 100 KB of shared memory per CTA displaces the serial kernels' staging
        // areas (measured: C2 -32 %, C5 at 8192 streams per GPU -13 %).
        const int tail_mode = env_int("AFSIM_TAIL", 0);
        const bool tail_dominated = !(a.structure & (ST_COMPRESSOR | ST_DEESSER));
        const bool use_tail = (a.structure & ST_LIMITER) && tail_supported(a.lookahead) && tail_mode != 1 &&
                              (tail_mode == 2 || (split && tail_dominated));
        if (use_tail) {
            AF_CUDA(h, sweep->mem.alloc(&a.st_lim, kStateLimiter * sp));
            AF_CUDA(h, sweep->mem.alloc(&a.tail_hist, static_cast<size_t>(kTailHistRows) * sp));
            AF_CUDA(h, tail_make_map(a.buf_a, a.ring_rows, S_pad, &batch->tail_map));
            batch->tail_err = sweep->d_tail_err;
        } else if (a.structure & ST_LIMITER) {
            AF_CUDA(h, alloc_ring(&a.buf_b, ring_elems));
            AF_CUDA(h, sweep->mem.alloc(&a.st_lim, kStateLimiter * sp));
            if (split) {
                AF_CUDA(h, alloc_ring(&a.buf_c, ring_elems));
                AF_CUDA(h, alloc_ring(&a.buf_p, ring_elems));
            } else {
                AF_CUDA(h, sweep->mem.alloc(&a.lim_sfx, static_cast<size_t>(a.lookahead + 1) * sp));
            }
        }
        a.stage_inputs = (split || auto_makeup) ? 1 : 0;  // the compressor's serial kernels only run staged
        {
            int n_w = 0;
            if (split && (a.structure & ST_LIMITER) && !use_tail) n_w = 1;
            if ((split || auto_makeup) && (a.structure & ST_COMPRESSOR)) n_w = 4;
            if (a.structure & ST_DEESSER) n_w = 13;  // the de-esser is always R/M split (afsim_deesser.h)
            for (int k = 0; k < n_w; ++k) AF_CUDA(h, alloc_ring(&a.w[k], ring_elems));
        }
        AF_CUDA(h, sweep->mem.alloc(&a.st_input, kStateInput * sp));
        AF_CUDA(h, sweep->mem.alloc(&a.st_tp, kStateTruePeak * sp));
        if (a.structure & ST_DEESSER) {
            AF_CUDA(h, sweep->mem.alloc(&a.st_deesser, kStateDeEsser * sp));
            double* de_tab = nullptr;
            AF_CUDA(h, sweep->mem.alloc(&de_tab, DE_FIELDS * sp));
            a.de_tab = de_tab;
        }
        if (a.structure & ST_COMPRESSOR) AF_CUDA(h, sweep->mem.alloc(&a.st_comp, kStateCompressor * sp));
        if (auto_makeup) {
            const MakeupConst mc = makeup_constants(fs, a.block_samples, T);
            if (a.block_samples / mc.slot > kMaxMakeupSub)
                return set_error(h, AFSIM_UNSUPPORTED, "auto makeup: block spans too many loudness-window slots");
            MakeupConst* d_mc = nullptr;
            AF_CUDA(h, sweep->mem.alloc(&d_mc, 1));
            AF_CUDA(h, cudaMemcpyAsync(d_mc, &mc, sizeof mc, cudaMemcpyHostToDevice, h->stream));
            AF_CUDA(h, cudaStreamSynchronize(h->stream));  // mc is a local
            a.mk_const = d_mc;
            batch->mk_ring_elems = static_cast<size_t>(2 * mc.n_slots + 2 * kMaxMakeupSub) * sp;
            AF_CUDA(h, sweep->mem.alloc(&a.mk_ring, batch->mk_ring_elems));
            AF_CUDA(h, sweep->mem.alloc(&a.st_mk, kStateMakeup * sp));
            if (opt.makeup_rows) AF_CUDA(h, sweep->mem.alloc(&a.mk_rows, static_cast<size_t>(3) * std::max(a.n_rows, 1) * sp));
            if (opt.vad) {
                std::vector<int64_t> voff(S_pad, -1);
                std::vector<double> pool;
                for (int s = 0; s < S; ++s) {
                    const double* v = opt.vad[members[s]];
                    if (!v) continue;
                    voff[s] = static_cast<int64_t>(pool.size());
                    pool.insert(pool.end(), v, v + a.n_rows);
                }
                double* d_vad = nullptr;
                int64_t* d_voff = nullptr;
                AF_CUDA(h, sweep->mem.alloc(&d_vad, pool.size()));
                AF_CUDA(h, sweep->mem.alloc(&d_voff, S_pad));
                if (!pool.empty())
                    AF_CUDA(h, cudaMemcpyAsync(d_vad, pool.data(), pool.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
                AF_CUDA(h, cudaMemcpyAsync(d_voff, voff.data(), S_pad * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
                AF_CUDA(h, cudaStreamSynchronize(h->stream));
                a.mk_vad = d_vad;
                a.mk_vad_off = d_voff;
            }
        }
        AF_CUDA(h, sweep->mem.alloc(&a.st_eq, static_cast<size_t>(kStateEqPerSection * kMaxSections) * sp));
        AF_CUDA(h, sweep->mem.alloc(&a.rows, static_cast<size_t>(4) * std::max(a.n_rows, 1) * sp));
        AF_CUDA(h, sweep->mem.alloc(&a.accum, sp));
        if (finalize_workspace_bytes(a.n_rows, a.n_pad) > kFinalizeSmemLimit)
            AF_CUDA(h, sweep->mem.alloc(&a.fin_scratch,
                                        static_cast<size_t>(S) * (finalize_workspace_bytes(a.n_rows, a.n_pad) / sizeof(float))));
        if (!sweep->d_accum_first) sweep->d_accum_first = a.accum;

        // stage list in chain order (block_processor.rs:106-161)
        batch->eq_k = S >= 32768 ? 10 : 5;
        batch->eq_k = env_int("AFSIM_EQ_K", batch->eq_k) == 10 ? 10 : 5;
        auto push_eq = [&]() {
            if (!(a.structure & ST_EQ)) return;
            for (uint32_t first = 0; first < max_sections; first += batch->eq_k)
                batch->stages.push_back({SK_EQ, static_cast<int>(first)});
        };
        bool shared_de_flag = false;  // set below, before the stage list is built
        // AFSIM_DE_CUT: 1 = never, 2 = always, else for the batches that take the R/M split kernels
        const int de_cut_mode = env_int("AFSIM_DE_CUT", 0);
        const bool de_cut = de_cut_mode == 2 || (de_cut_mode != 1 && split);
        auto push_deesser = [&]() {
            if (!(a.structure & ST_DEESSER)) return;
            for (int op : {SP_DE_RA, SP_DE_MB, SP_DE_RC, SP_DE_MC2, SP_DE_RC3}) {
                if (shared_de_flag && (op == SP_DE_RA || op == SP_DE_MB)) continue;
                if (op == SP_DE_RC && de_cut) {  // few streams: R_c1 as two short serial kernels around a map
                    for (int cut_op : {SP_DE_RC1A, SP_DE_MC1B, SP_DE_RC1C}) batch->stages.push_back({SK_SPLIT, cut_op});
                    continue;
                }
                batch->stages.push_back({SK_SPLIT, op});
            }
        };
        // Shared prefix: when several streams of the batch read the same passage (a candidate sweep), the input stage
        // -- identical for all of them -- runs once per distinct passage and a copy kernel fans it out.  When the EQ
        // is the first stage after it and the streams of a passage also share their EQ (a compressor grid over one
        // EQ setting), the EQ runs on the distinct (passage, EQ) pairs as well and its output is what is fanned out.
        bool shared_eq = false, shared_front = false, shared_de = false;
        uint32_t shared_max_sections = 0;
        {
            const bool eq_first = (a.structure & ST_EQ) && !(a.structure & ST_INPUT_TRUE_PEAK) &&
                                  (!(a.structure & ST_DEESSER) || (a.structure & ST_EQ_BEFORE_DEESSER));
            // EQ class of every candidate used by the batch: same sections, coefficients and fade flag
            std::map<std::vector<unsigned char>, uint32_t> eq_classes;
            std::vector<uint32_t> eq_class_of(plans.size(), 0);
            if (eq_first) {
                std::vector<bool> classified(plans.size(), false);
                for (int s = 0; s < S; ++s) {
                    if (classified[cand[s]]) continue;  // once per candidate, not per stream
                    classified[cand[s]] = true;
                    eq_class_of[cand[s]] = eq_classes.emplace(eq_class_key(plans[cand[s]].params), static_cast<uint32_t>(eq_classes.size())).first->second;
                }
            }
            // de-esser first: its detector front (R_a + M_b) depends on the band split and the fixed detector time
            // constants only -- classify those
            const bool de_first = (a.structure & ST_DEESSER) && !(a.structure & ST_EQ_BEFORE_DEESSER) &&
                                  !(a.structure & ST_INPUT_TRUE_PEAK);
            if (de_first) {
                std::map<std::vector<unsigned char>, uint32_t> de_classes;
                std::vector<bool> classified(plans.size(), false);
                for (int s = 0; s < S; ++s) {
                    if (classified[cand[s]]) continue;
                    classified[cand[s]] = true;
                    const CandidateParams& p = plans[cand[s]].params;
                    std::vector<unsigned char> key(30 * 8 + 2 * 8 + sizeof p.de_det0);
                    std::memcpy(key.data(), p.de + DE_DET, 30 * 8);
                    std::memcpy(key.data() + 240, p.de + DE_DET_ATTACK, 8);
                    std::memcpy(key.data() + 248, p.de + DE_DET_RELEASE, 8);
                    std::memcpy(key.data() + 256, p.de_det0, sizeof p.de_det0);
                    eq_class_of[cand[s]] = de_classes.emplace(std::move(key), static_cast<uint32_t>(de_classes.size())).first->second;
                }
            }
            auto distinct_streams = [&](bool with_eq, std::vector<uint32_t>& uidx, std::vector<uint64_t>& usrc,
                                        std::vector<uint32_t>& ucand) {
                std::map<std::pair<uint64_t, uint32_t>, uint32_t> distinct;
                uidx.assign(S_pad, 0);
                usrc.clear();
                ucand.clear();
                for (int s = 0; s < S; ++s) {
                    const std::pair<uint64_t, uint32_t> key(src_off[s], with_eq ? eq_class_of[cand[s]] : 0u);
                    auto it = distinct.find(key);
                    if (it == distinct.end()) {
                        it = distinct.emplace(key, static_cast<uint32_t>(usrc.size())).first;
                        usrc.push_back(src_off[s]);
                        ucand.push_back(cand[s]);
                    }
                    uidx[s] = it->second;
                }
            };
            std::vector<uint32_t> uidx, ucand;
            std::vector<uint64_t> usrc;
            const int share_mode = env_int("AFSIM_SHARED_INPUT", 1);  // 1: input (+ EQ), 2: off, 3: input only
            if (eq_first && share_mode == 1) {
                distinct_streams(true, uidx, usrc, ucand);
                shared_eq = S >= 64 && static_cast<int>(usrc.size()) * 4 <= S;
            }
            if (de_first && share_mode == 1) {
                distinct_streams(true, uidx, usrc, ucand);
                shared_de = S >= 64 && static_cast<int>(usrc.size()) * 4 <= S;
            }
            if (!shared_eq && !shared_de) distinct_streams(false, uidx, usrc, ucand);
            const int U = static_cast<int>(usrc.size());
            if (S >= 64 && U * 4 <= S && share_mode != 2) {
                const int U_pad = round_up(U, 32);
                for (int u = 0; u < U; ++u) shared_max_sections = std::max(shared_max_sections, plans[ucand[u]].params.n_sections);
                usrc.resize(U_pad, 0);
                ucand.resize(U_pad, 0);
                BatchArgs& ua = batch->shared_input;
                ua = a;
                ua.n_streams = U;
                ua.stride = U_pad;
                uint32_t *d_uidx = nullptr, *d_ucand = nullptr;
                uint64_t* d_usrc = nullptr;
                float *d_ubuf = nullptr, *d_urows = nullptr;
                StreamAccum* d_uaccum = nullptr;
                double* d_ustate = nullptr;
                batch->shared_rows_elems = static_cast<size_t>(std::max(a.n_rows, 1)) * U_pad;
                AF_CUDA(h, sweep->mem.alloc(&d_uidx, S_pad));
                AF_CUDA(h, sweep->mem.alloc(&d_ucand, U_pad));
                AF_CUDA(h, sweep->mem.alloc(&d_usrc, U_pad));
                AF_CUDA(h, alloc_ring(&d_ubuf, static_cast<size_t>(a.ring_rows) * U_pad));
                AF_CUDA(h, sweep->mem.alloc(&d_urows, batch->shared_rows_elems));
                AF_CUDA(h, sweep->mem.alloc(&d_uaccum, U_pad));
                AF_CUDA(h, sweep->mem.alloc(&d_ustate, static_cast<size_t>(kStateInput) * U_pad));
                if (shared_eq) AF_CUDA(h, sweep->mem.alloc(&ua.st_eq, static_cast<size_t>(kStateEqPerSection * kMaxSections) * U_pad));
                // the compressor front as well, when the compressor follows the EQ directly (the makeup stage R7 works in
                // place on the stream's own signal: auto-makeup batches keep their own front)
                shared_front = shared_eq && !auto_makeup && (a.structure & ST_COMPRESSOR) && !(a.structure & ST_DEESSER) &&
                               share_mode == 1;
                if (shared_front) {
                    for (int k = 0; k < 4; ++k) AF_CUDA(h, alloc_ring(&ua.w[k], static_cast<size_t>(a.ring_rows) * U_pad));
                    AF_CUDA(h, sweep->mem.alloc(&ua.st_comp, static_cast<size_t>(kStateCompressor) * U_pad));
                }
                if (shared_de) {  // detector front of the de-esser on the distinct (passage, detector) pairs
                    for (int k = 0; k < 7; ++k) AF_CUDA(h, sweep->mem.alloc(&ua.w[k], static_cast<size_t>(a.ring_rows) * U_pad));
                    AF_CUDA(h, sweep->mem.alloc(&ua.st_deesser, static_cast<size_t>(kStateDeEsser) * U_pad));
                    double* u_de_tab = nullptr;
                    AF_CUDA(h, sweep->mem.alloc(&u_de_tab, static_cast<size_t>(DE_FIELDS) * U_pad));
                    ua.de_tab = u_de_tab;
                }
                std::vector<double> utab(static_cast<size_t>(MT_FIELDS) * U_pad);
                for (int u = 0; u < U_pad; ++u) fill_map_tab(utab.data(), U_pad, u, plans[ucand[u]].params);
                double* d_utab = nullptr;
                AF_CUDA(h, sweep->mem.alloc(&d_utab, utab.size()));
                AF_CUDA(h, cudaMemcpyAsync(d_utab, utab.data(), utab.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
                ua.map_tab = d_utab;
                AF_CUDA(h, cudaMemcpyAsync(d_uidx, uidx.data(), S_pad * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
                AF_CUDA(h, cudaMemcpyAsync(d_ucand, ucand.data(), U_pad * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
                AF_CUDA(h, cudaMemcpyAsync(d_usrc, usrc.data(), U_pad * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
                AF_CUDA(h, cudaStreamSynchronize(h->stream));
                ua.cand = d_ucand;
                ua.src_off = d_usrc;
                ua.buf_a = d_ubuf;
                ua.rows = d_urows;
                ua.accum = d_uaccum;
                ua.st_input = d_ustate;
                ua.stage_inputs = (shared_front || shared_de) ? 1 : 0;  // the shared serial kernels run staged
                if (shared_front) {
                    a.in_det = ua.w[0];
                    a.in_wdb = ua.w[1];
                    a.in_ipk = ua.w[2];
                }
                if (shared_de)
                    for (int k = 0; k < 7; ++k) a.in_de[k] = ua.w[k];
                a.in_unique = d_uidx;
                a.in_src = d_ubuf;
                a.in_rows = d_urows;
                a.in_accum = d_uaccum;
                a.in_stride = U_pad;
            } else {
                shared_eq = false;
                shared_de = false;
            }
        }
        if (!shared_eq) shared_front = false;
        batch->shared_deesser = shared_de;
        shared_de_flag = shared_de;
        if (batch->shared_input.n_streams > 0) {
            batch->stages.push_back({SK_INPUT_SHARED, 0});
            if (shared_eq)
                for (uint32_t first = 0; first < shared_max_sections; first += batch->eq_k)
                    batch->stages.push_back({SK_EQ_SHARED, static_cast<int>(first)});
            if (shared_front) {
                batch->stages.push_back({SK_SPLIT_SHARED, SP_COMP_R1});
                batch->stages.push_back({SK_SPLIT_SHARED, SP_COMP_M2});
            }
            if (shared_de) {
                batch->stages.push_back({SK_SPLIT_SHARED, SP_DE_RA});
                batch->stages.push_back({SK_SPLIT_SHARED, SP_DE_MB});
            }
            batch->stages.push_back({SK_INPUT_FANOUT, 0});
        } else {
            batch->stages.push_back({SK_INPUT, 0});
        }
        if (a.structure & ST_INPUT_TRUE_PEAK) batch->stages.push_back({SK_INPUT_TP, 0});
        if (shared_eq) {
            push_deesser();  // the EQ already ran on the distinct (passage, EQ) pairs
        } else if (a.structure & ST_EQ_BEFORE_DEESSER) {
            push_eq();
            push_deesser();
        } else {
            push_deesser();
            push_eq();
        }
        if (a.structure & ST_COMPRESSOR) {
            if (split || auto_makeup) {
                for (int op : {SP_COMP_R1, SP_COMP_M2, SP_COMP_R3, SP_COMP_M4, SP_COMP_R5, SP_COMP_M6})
                    if (!(shared_front && (op == SP_COMP_R1 || op == SP_COMP_M2)))  // rendered once per (passage, EQ) pair
                        batch->stages.push_back({SK_SPLIT, op});
                if (auto_makeup) batch->stages.push_back({SK_SPLIT, SP_COMP_R7});
            } else {
                batch->stages.push_back({SK_COMPRESSOR, 0});
            }
        }
        if (a.structure & ST_LIMITER) {
            if (use_tail)
                batch->stages.push_back({SK_TAIL, 0});
            else if (split)
                for (int op : {SP_LIM_M, SP_LIM_R, SP_TP_FIR_IN, SP_TP_R, SP_TP_FIR_OUT}) batch->stages.push_back({SK_SPLIT, op});
            else
                batch->stages.push_back({SK_LIMITER, 0});
        }
        if (!((split || use_tail) && (a.structure & ST_LIMITER))) batch->stages.push_back({SK_OUTPUT, 0});
        if (static_cast<int>(batch->stages.size()) > kMaxStages) return set_error(h, AFSIM_UNSUPPORTED, "too many stages");
        batch->events.resize(batch->stages.size() * static_cast<size_t>(slots));
        for (cudaEvent_t& e : batch->events) AF_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        sweep->kernels_per_launch += static_cast<int>(batch->stages.size()) * n_chunks + 1 + ((a.structure & ST_DEESSER) ? 1 : 0) + (shared_de ? 1 : 0);
        sweep->batches.push_back(std::move(batch));
    }
    AF_CUDA(h, cudaEventCreate(&sweep->ev_start));
    AF_CUDA(h, cudaEventCreate(&sweep->ev_stop));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    sweep->owner = h;
    h->live.insert(sweep.get());
    *out_sweep = sweep.release();
    return AFSIM_OK;
}

cudaError_t launch_stage(const Batch& b, const StageDesc& st, const ChunkArgs& ck, cudaStream_t stream) {
    switch (st.kind) {
        case SK_INPUT: return launch_input(b.args, ck, stream);
        case SK_INPUT_SHARED: return launch_input(b.shared_input, ck, stream);
        case SK_INPUT_FANOUT: return launch_input_fanout(b.args, ck, stream);
        case SK_EQ_SHARED: return launch_eq(b.shared_input, ck, st.arg, b.eq_k, stream);
        case SK_SPLIT_SHARED: return launch_split(static_cast<SplitOp>(st.arg), b.shared_input, ck, stream);
        case SK_INPUT_TP: return launch_input_true_peak(b.args, ck, stream);
        case SK_EQ: return launch_eq(b.args, ck, st.arg, b.eq_k, stream);
        case SK_COMPRESSOR: return launch_compressor(b.args, ck, stream);
        case SK_LIMITER: return launch_limiter(b.args, ck, stream);
        case SK_SPLIT: return launch_split(static_cast<SplitOp>(st.arg), b.args, ck, stream);
        case SK_TAIL: return launch_tail(b.args, ck, b.tail_map, b.tail_err, stream);
        default: return launch_output(b.args, ck, (b.args.structure & ST_LIMITER) != 0, stream);
    }
}

// One batch: the chunk x stage wavefront.  Stage i runs on its own CUDA stream (so its chunks stay
// ordered, which carries the parked state); stage i of chunk c waits for stage i-1 of chunk c; the
// first stage of chunk c waits for the last stage of chunk c - slots + 1 before it reuses a ring slot.
struct WavefrontTrace {  // timing events around the launches of chunks [c0, c0 + n) in the live wavefront
    int c0 = 0, n = 0;
    std::vector<cudaEvent_t> ev;  // [chunk][stage][2]
};

int run_batch(AfsimHandle* h, Batch& b, WavefrontTrace* trace = nullptr) {
    const BatchArgs& a = b.args;
    const int n_stages = static_cast<int>(b.stages.size());
    const int T = a.n_samples;
    AF_CUDA(h, cudaMemsetAsync(a.accum, 0, static_cast<size_t>(a.stride) * sizeof(StreamAccum), h->stream));
    AF_CUDA(h, cudaMemsetAsync(a.rows, 0, static_cast<size_t>(4) * std::max(a.n_rows, 1) * a.stride * sizeof(float), h->stream));
    if (b.shared_input.n_streams > 0) {
        AF_CUDA(h, cudaMemsetAsync(b.shared_input.accum, 0, static_cast<size_t>(b.shared_input.stride) * sizeof(StreamAccum), h->stream));
        AF_CUDA(h, cudaMemsetAsync(b.shared_input.rows, 0, b.shared_rows_elems * sizeof(float), h->stream));
    }
    if (a.structure & ST_DEESSER) AF_CUDA(h, launch_expand_deesser(a, h->stream));
    if (b.shared_deesser) AF_CUDA(h, launch_expand_deesser(b.shared_input, h->stream));
    if (a.mk_ring) AF_CUDA(h, cudaMemsetAsync(a.mk_ring, 0, b.mk_ring_elems * sizeof(double), h->stream));
    if (a.mk_rows) AF_CUDA(h, cudaMemsetAsync(a.mk_rows, 0, static_cast<size_t>(3) * std::max(a.n_rows, 1) * a.stride * sizeof(float), h->stream));
    if (T > 0) {
        AF_CUDA(h, cudaEventRecord(h->ev_fork, h->stream));
        auto stream_of = [&](int i) {
            const StageDesc& sd = b.stages[i];
            if (sd.kind == SK_INPUT_FANOUT) return h->stage_stream_map[i];
            const bool is_map = (sd.kind == SK_SPLIT || sd.kind == SK_SPLIT_SHARED) && (sd.arg == SP_COMP_M2 || sd.arg == SP_COMP_M4 || sd.arg == SP_COMP_M6 ||
                                                        sd.arg == SP_LIM_M || sd.arg == SP_TP_FIR_IN || sd.arg == SP_TP_FIR_OUT ||
                                                        sd.arg == SP_DE_MB || sd.arg == SP_DE_MC2 || sd.arg == SP_DE_MC1B);
            return is_map ? h->stage_stream_map[i] : h->stage_stream[i];
        };
        for (int i = 0; i < n_stages; ++i) AF_CUDA(h, cudaStreamWaitEvent(stream_of(i), h->ev_fork, 0));
        const int n_chunks = (T + b.chunk - 1) / b.chunk;
        for (int c = 0; c < n_chunks; ++c) {
            ChunkArgs ck;
            ck.n0 = c * b.chunk;
            ck.len = std::min(b.chunk, T - ck.n0);
            const int slot = c % b.slots;
            ck.row0 = slot * b.chunk;
            for (int i = 0; i < n_stages; ++i) {
                cudaStream_t st = stream_of(i);
                if (i > 0) AF_CUDA(h, cudaStreamWaitEvent(st, b.events[static_cast<size_t>(i - 1) * b.slots + slot], 0));
                if (i == 0 && c - b.slots + 1 >= 0) {
                    const int old_slot = (c - b.slots + 1) % b.slots;
                    AF_CUDA(h, cudaStreamWaitEvent(st, b.events[static_cast<size_t>(n_stages - 1) * b.slots + old_slot], 0));
                }
                const bool traced = trace && c >= trace->c0 && c < trace->c0 + trace->n;
                const size_t te = traced ? (static_cast<size_t>(c - trace->c0) * n_stages + i) * 2 : 0;
                if (traced) AF_CUDA(h, cudaEventRecord(trace->ev[te], st));  // after the waits: the launch is eligible
                AF_CUDA(h, launch_stage(b, b.stages[i], ck, st));
                if (traced) AF_CUDA(h, cudaEventRecord(trace->ev[te + 1], st));
                AF_CUDA(h, cudaEventRecord(b.events[static_cast<size_t>(i) * b.slots + slot], st));
            }
        }
        // join: the last stage's stream has seen every other stage through the event chain
        const int last_slot = (n_chunks - 1) % b.slots;
        AF_CUDA(h, cudaStreamWaitEvent(h->stream, b.events[static_cast<size_t>(n_stages - 1) * b.slots + last_slot], 0));
    }
    AF_CUDA(h, launch_finalize(a, h->stream));
    return AFSIM_OK;
}

}  // namespace

// =====================================================================================================================
extern "C" {

int afsim_abi_version(void) { return AFSIM_ABI_VERSION; }

const char* afsim_create_error(void) { return g_create_error.c_str(); }

int afsim_create(int device_ordinal, void* cuda_stream, AfsimHandle** out_handle) {
    g_create_error.clear();
    if (!out_handle) {
        g_create_error = "out_handle is null";
        return AFSIM_INVALID_ARGUMENT;
    }
    *out_handle = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0") +
                         " (libafsim has no CPU path)";
        cudaGetLastError();
        return AFSIM_CUDA_ERROR;
    }
    if (device_ordinal < 0 || device_ordinal >= count) {
        g_create_error = "device ordinal out of range";
        return AFSIM_INVALID_ARGUMENT;
    }
    cudaDeviceProp prop;
    err = cudaGetDeviceProperties(&prop, device_ordinal);
    if (err != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(err);
        return AFSIM_CUDA_ERROR;
    }
    if (prop.major != 10) {
        g_create_error = "libafsim is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        return AFSIM_CUDA_ERROR;
    }
    err = cudaSetDevice(device_ordinal);
    if (err != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(err);
        return AFSIM_CUDA_ERROR;
    }
    err = configure_kernels();
    if (err == cudaSuccess) err = tail_configure();
    if (err != cudaSuccess) {
        g_create_error = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err);
        return AFSIM_CUDA_ERROR;
    }
    AfsimHandle* h = new AfsimHandle();
    h->device = device_ordinal;
    h->pool.max_cached = static_cast<size_t>(env_int("AFSIM_POOL_MAX_GB", 96)) << 30;
    auto fail = [&](const char* what, cudaError_t e) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(e);
        afsim_destroy(h);  // destroys whatever was created so far
        return AFSIM_CUDA_ERROR;
    };
    if (cuda_stream) {
        h->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        err = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (err != cudaSuccess) return fail("cudaStreamCreate", err);
        h->own_stream = true;
    }
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    for (int i = 0; i < kMaxStages; ++i) {
        err = cudaStreamCreateWithPriority(&h->stage_stream[i], cudaStreamNonBlocking, prio_greatest);
        if (err == cudaSuccess) err = cudaStreamCreateWithPriority(&h->stage_stream_map[i], cudaStreamNonBlocking, prio_least);
        if (err != cudaSuccess) return fail("cudaStreamCreate", err);
    }
    err = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (err != cudaSuccess) return fail("cudaEventCreate", err);
    *out_handle = h;
    return AFSIM_OK;
}

void afsim_destroy(AfsimHandle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (!h->live.empty()) {
        // sweeps the caller still holds outlive the handle: they keep their device memory (freed with cudaFree by
        // afsim_sweep_release(NULL, sweep)) but no longer point into this handle's pool
        cudaDeviceSynchronize();
        for (AfsimSweep* sw : h->live) {
            sw->mem.pool = nullptr;
            sw->owner = nullptr;
            sw->quiesced = true;
        }
        h->live.clear();
    }
    for (int i = 0; i < kMaxStages; ++i) {
        if (h->stage_stream[i]) cudaStreamDestroy(h->stage_stream[i]);
        if (h->stage_stream_map[i]) cudaStreamDestroy(h->stage_stream_map[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    h->resample.drop();
    h->resample.drop_staging();
    h->pool.trim();
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int afsim_trim(AfsimHandle* h) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    std::lock_guard<std::recursive_mutex> lock(h->call_mu);
    cudaSetDevice(h->device);
    h->pool.trim();
    return AFSIM_OK;
}

const char* afsim_last_error(const AfsimHandle* h) { return h ? h->error.c_str() : "null handle"; }

void afsim_chain_settings_default(AfChainSettings* out) {
    if (out) chain_settings_default(out);
}
void afsim_default_bands(AfBand out[AFSIM_NUM_BANDS]) {
    if (out) default_bands(out);
}

// ---- sweeps ----------------------------------------------------------------------------------------------------

int afsim_sweep_prepare(AfsimHandle* h, const float* const* passages, const size_t* passage_len, size_t n_passages,
                        double sample_rate, const AfCandidate* candidates, size_t n_candidates,
                        const uint32_t* pair_passage, const uint32_t* pair_candidate, size_t n_pairs, int want_audio,
                        AfsimSweep** out_sweep) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    if ((n_passages && (!passages || !passage_len)) || (n_candidates && !candidates))
        return set_error(h, AFSIM_INVALID_ARGUMENT, "null passages / candidates");
    PassageSource src;
    src.host = passages;
    return build_sweep(h, src, passage_len, n_passages, sample_rate, candidates, n_candidates, nullptr, pair_passage,
                       pair_candidate, n_pairs, want_audio, out_sweep);
}

int afsim_sweep_prepare_synthetic(AfsimHandle* h, int kind, size_t n_passages, size_t passage_len, double sample_rate,
                                  const AfCandidate* candidates, size_t n_candidates, const uint32_t* pair_passage,
                                  const uint32_t* pair_candidate, size_t n_pairs, int want_audio, AfsimSweep** out_sweep) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    if (n_candidates && !candidates) return set_error(h, AFSIM_INVALID_ARGUMENT, "null candidates");
    if (kind != 0 && kind != 1) return set_error(h, AFSIM_INVALID_ARGUMENT, "synthetic kind must be 0 or 1");
    std::vector<size_t> lens(n_passages, passage_len);
    PassageSource src;
    src.host = nullptr;
    src.synth_kind = kind;
    return build_sweep(h, src, lens.data(), n_passages, sample_rate, candidates, n_candidates, nullptr, pair_passage,
                       pair_candidate, n_pairs, want_audio, out_sweep);
}

int afsim_sweep_launch(AfsimHandle* h, AfsimSweep* sweep) {
    if (!h || !sweep) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    AF_CUDA(h, cudaSetDevice(h->device));
    AF_CUDA(h, cudaEventRecord(sweep->ev_start, h->stream));
    for (auto& b : sweep->batches) {
        const int rc = run_batch(h, *b);
        if (rc != AFSIM_OK) return rc;
    }
    AF_CUDA(h, cudaEventRecord(sweep->ev_stop, h->stream));
    sweep->launched = true;
    return AFSIM_OK;
}

int afsim_sweep_collect(AfsimHandle* h, AfsimSweep* sweep, AfChainMetrics* out_metrics) {
    if (!h || !sweep || (!out_metrics && sweep->n_pairs)) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    AF_CUDA(h, cudaSetDevice(h->device));
    if (sweep->n_pairs)
        AF_CUDA(h, cudaMemcpyAsync(out_metrics, sweep->d_metrics, sweep->n_pairs * sizeof(AfChainMetrics),
                                   cudaMemcpyDeviceToHost, h->stream));
    int tail_err = 0;
    AF_CUDA(h, cudaMemcpyAsync(&tail_err, sweep->d_tail_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (tail_err) return set_error(h, AFSIM_CUDA_ERROR, "fused tail kernel: pipeline watchdog fired (results invalid)");
    float ms = 0.0f;
    if (sweep->launched && cudaEventElapsedTime(&ms, sweep->ev_start, sweep->ev_stop) == cudaSuccess && sweep->n_pairs)
        for (size_t i = 0; i < sweep->n_pairs; ++i) out_metrics[i].candidate_runtime_ms = static_cast<double>(ms) / sweep->n_pairs;
    return AFSIM_OK;
}

int afsim_sweep_collect_audio(AfsimHandle* h, AfsimSweep* sweep, size_t pair, float* out_audio, size_t n) {
    if (!h || !sweep) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    if (!sweep->d_audio) return set_error(h, AFSIM_INVALID_ARGUMENT, "sweep was prepared without want_audio");
    if (pair >= sweep->n_pairs || n > sweep->pair_len[pair] || (!out_audio && n))
        return set_error(h, AFSIM_INVALID_ARGUMENT, "audio request out of range");
    AF_CUDA(h, cudaSetDevice(h->device));
    if (n)
        AF_CUDA(h, cudaMemcpyAsync(out_audio, sweep->d_audio + sweep->audio_off[pair], n * sizeof(float),
                                   cudaMemcpyDeviceToHost, h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    return AFSIM_OK;
}

int afsim_sweep_status(AfsimHandle* h, AfsimSweep* sweep) {
    if (!h || !sweep) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    AF_CUDA(h, cudaSetDevice(h->device));
    int tail_err = 0;
    AF_CUDA(h, cudaMemcpyAsync(&tail_err, sweep->d_tail_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (tail_err) return set_error(h, AFSIM_CUDA_ERROR, "fused tail kernel: pipeline watchdog fired (results invalid)");
    return AFSIM_OK;
}

void* afsim_sweep_metrics_device_ptr(AfsimSweep* sweep) { return sweep ? sweep->d_metrics : nullptr; }

int afsim_sweep_kernel_count(const AfsimSweep* sweep) { return sweep ? sweep->kernels_per_launch : 0; }

int afsim_sweep_last_render_ms(AfsimHandle* h, AfsimSweep* sweep, float* out_ms) {
    if (!h || !sweep || !out_ms) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    if (!sweep->launched) return set_error(h, AFSIM_INVALID_ARGUMENT, "sweep has not been launched");
    AF_CUDA(h, cudaSetDevice(h->device));
    AF_CUDA(h, cudaEventSynchronize(sweep->ev_stop));
    AF_CUDA(h, cudaEventElapsedTime(out_ms, sweep->ev_start, sweep->ev_stop));
    return AFSIM_OK;
}

void afsim_sweep_release(AfsimHandle* h, AfsimSweep* sweep) {
    if (!sweep) return;
    h = sweep->owner;  // the handle the sweep was prepared on, or nullptr once that handle was destroyed
    if (h) {
        std::lock_guard<std::recursive_mutex> lock(h->call_mu);
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
        for (int i = 0; i < kMaxStages; ++i) {
            cudaStreamSynchronize(h->stage_stream[i]);
            cudaStreamSynchronize(h->stage_stream_map[i]);
        }
        sweep->quiesced = true;
        h->live.erase(sweep);
        delete sweep;
        return;
    }
    cudaSetDevice(sweep->device);
    delete sweep;
}

int afsim_sweep_profile_stages(AfsimHandle* h, AfsimSweep* sweep, int max_chunks, int capacity, int* out_kind,
                               float* out_ms, int* out_launches, int* out_n) {
    if (!h || !sweep || !out_kind || !out_ms || !out_launches || !out_n) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    *out_n = 0;
    if (sweep->batches.empty()) return AFSIM_OK;
    AF_CUDA(h, cudaSetDevice(h->device));
    Batch& b = *sweep->batches[0];
    const BatchArgs& a = b.args;
    const int n_stages = static_cast<int>(b.stages.size());
    if (capacity < n_stages + 1) return set_error(h, AFSIM_INVALID_ARGUMENT, "capacity too small");
    const int T = a.n_samples;
    const int n_chunks = T > 0 ? (T + b.chunk - 1) / b.chunk : 0;
    const int timed_chunks = max_chunks > 0 ? std::min(max_chunks, n_chunks) : n_chunks;
    std::vector<cudaEvent_t> ev(static_cast<size_t>(timed_chunks) * n_stages * 2 + 2);
    for (cudaEvent_t& e : ev) AF_CUDA(h, cudaEventCreate(&e));
    auto destroy = [&]() {
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
    };
    cudaError_t err = cudaMemsetAsync(a.accum, 0, static_cast<size_t>(a.stride) * sizeof(StreamAccum), h->stream);
    if (err == cudaSuccess)
        err = cudaMemsetAsync(a.rows, 0, static_cast<size_t>(4) * std::max(a.n_rows, 1) * a.stride * sizeof(float), h->stream);
    if (err == cudaSuccess && (a.structure & ST_DEESSER)) err = launch_expand_deesser(a, h->stream);
    if (err == cudaSuccess && b.shared_deesser) err = launch_expand_deesser(b.shared_input, h->stream);
    if (err == cudaSuccess && a.mk_ring) err = cudaMemsetAsync(a.mk_ring, 0, b.mk_ring_elems * sizeof(double), h->stream);
    for (int c = 0; c < n_chunks && err == cudaSuccess; ++c) {
        ChunkArgs ck;
        ck.n0 = c * b.chunk;
        ck.len = std::min(b.chunk, T - ck.n0);
        ck.row0 = (c % b.slots) * b.chunk;
        for (int i = 0; i < n_stages && err == cudaSuccess; ++i) {
            const bool timed = c < timed_chunks;
            const size_t e0 = (static_cast<size_t>(c) * n_stages + i) * 2;
            if (timed) err = cudaEventRecord(ev[e0], h->stream);
            if (err == cudaSuccess) err = launch_stage(b, b.stages[i], ck, h->stream);
            if (timed && err == cudaSuccess) err = cudaEventRecord(ev[e0 + 1], h->stream);
        }
    }
    const size_t fin = ev.size() - 2;
    if (err == cudaSuccess) err = cudaEventRecord(ev[fin], h->stream);
    if (err == cudaSuccess) err = launch_finalize(a, h->stream);
    if (err == cudaSuccess) err = cudaEventRecord(ev[fin + 1], h->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(h->stream);
    if (err != cudaSuccess) {
        destroy();
        return cuda_fail(h, err, "afsim_sweep_profile_stages");
    }
    static const int kind_map[] = {AF_STAGE_INPUT, AF_STAGE_INPUT_TRUE_PEAK, AF_STAGE_DEESSER, AF_STAGE_EQ,
                                   AF_STAGE_COMPRESSOR, AF_STAGE_LIMITER, AF_STAGE_OUTPUT, 0, AF_STAGE_INPUT, AF_STAGE_INPUT_FANOUT, AF_STAGE_EQ, 0, AF_STAGE_TAIL};
    for (int i = 0; i < n_stages; ++i) {
        double total = 0.0;
        for (int c = 0; c < timed_chunks; ++c) {
            float ms = 0.0f;
            const size_t e0 = (static_cast<size_t>(c) * n_stages + i) * 2;
            cudaEventElapsedTime(&ms, ev[e0], ev[e0 + 1]);
            total += ms;
        }
        out_kind[i] = (b.stages[i].kind == SK_SPLIT || b.stages[i].kind == SK_SPLIT_SHARED) ? AF_STAGE_SPLIT_BASE + b.stages[i].arg : kind_map[b.stages[i].kind];
        out_ms[i] = static_cast<float>(total);
        out_launches[i] = timed_chunks;
    }
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ev[fin], ev[fin + 1]);
    out_kind[n_stages] = AF_STAGE_FINALIZE;
    out_ms[n_stages] = ms;
    out_launches[n_stages] = 1;
    *out_n = n_stages + 1;
    destroy();
    return AFSIM_OK;
}

int afsim_sweep_batch_info(const AfsimSweep* sweep, int capacity, int out_info[6], int* out_stage_streams, int* out_n) {
    if (!sweep || !out_info || !out_stage_streams || !out_n) return AFSIM_INVALID_ARGUMENT;
    *out_n = 0;
    for (int k = 0; k < 6; ++k) out_info[k] = 0;
    out_info[0] = static_cast<int>(sweep->batches.size());
    if (sweep->batches.empty()) return AFSIM_OK;
    const Batch& b = *sweep->batches[0];
    const int n_stages = static_cast<int>(b.stages.size());
    if (capacity < n_stages) return AFSIM_INVALID_ARGUMENT;
    out_info[1] = b.args.n_streams;
    out_info[2] = b.chunk;
    out_info[3] = b.slots;
    out_info[4] = n_stages;
    out_info[5] = b.args.n_samples;
    for (int i = 0; i < n_stages; ++i) {
        const StageKind k = b.stages[i].kind;
        const bool shared = k == SK_INPUT_SHARED || k == SK_EQ_SHARED || k == SK_SPLIT_SHARED;
        out_stage_streams[i] = shared ? b.shared_input.n_streams : b.args.n_streams;
    }
    *out_n = n_stages;
    return AFSIM_OK;
}

int afsim_sweep_profile_wavefront(AfsimHandle* h, AfsimSweep* sweep, int first_chunk, int n_chunks, int capacity, int* out_kind,
                                  float* out_busy_ms, float* out_period_ms, int* out_n) {
    if (!h || !sweep || !out_kind || !out_busy_ms || !out_period_ms || !out_n) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    *out_n = 0;
    if (sweep->batches.empty()) return AFSIM_OK;
    AF_CUDA(h, cudaSetDevice(h->device));
    Batch& b = *sweep->batches[0];
    const int n_stages = static_cast<int>(b.stages.size());
    if (capacity < n_stages) return set_error(h, AFSIM_INVALID_ARGUMENT, "capacity too small");
    const int total_chunks = b.args.n_samples > 0 ? (b.args.n_samples + b.chunk - 1) / b.chunk : 0;
    WavefrontTrace tr;
    tr.c0 = std::max(0, std::min(first_chunk, total_chunks - 1));
    tr.n = std::max(0, std::min(n_chunks, total_chunks - tr.c0));
    if (tr.n < 2) return set_error(h, AFSIM_INVALID_ARGUMENT, "need at least two chunks to trace");
    tr.ev.resize(static_cast<size_t>(tr.n) * n_stages * 2);
    for (cudaEvent_t& e : tr.ev) AF_CUDA(h, cudaEventCreate(&e));
    int rc = run_batch(h, b, &tr);
    cudaError_t err = cudaStreamSynchronize(h->stream);
    if (rc == AFSIM_OK && err != cudaSuccess) rc = cuda_fail(h, err, "afsim_sweep_profile_wavefront");
    if (rc == AFSIM_OK) {
        static const int kind_map[] = {AF_STAGE_INPUT, AF_STAGE_INPUT_TRUE_PEAK, AF_STAGE_DEESSER, AF_STAGE_EQ,
                                       AF_STAGE_COMPRESSOR, AF_STAGE_LIMITER, AF_STAGE_OUTPUT, 0, AF_STAGE_INPUT, AF_STAGE_INPUT_FANOUT, AF_STAGE_EQ, 0, AF_STAGE_TAIL};
        for (int i = 0; i < n_stages; ++i) {
            double busy = 0.0;
            for (int c = 0; c < tr.n; ++c) {
                float ms = 0.0f;
                const size_t e0 = (static_cast<size_t>(c) * n_stages + i) * 2;
                cudaEventElapsedTime(&ms, tr.ev[e0], tr.ev[e0 + 1]);
                busy += ms;
            }
            float span = 0.0f;  // completion of the first traced chunk -> completion of the last one, on this stage
            cudaEventElapsedTime(&span, tr.ev[static_cast<size_t>(i) * 2 + 1],
                                 tr.ev[(static_cast<size_t>(tr.n - 1) * n_stages + i) * 2 + 1]);
            out_kind[i] = (b.stages[i].kind == SK_SPLIT || b.stages[i].kind == SK_SPLIT_SHARED) ? AF_STAGE_SPLIT_BASE + b.stages[i].arg : kind_map[b.stages[i].kind];
            out_busy_ms[i] = static_cast<float>(busy / tr.n);
            out_period_ms[i] = span / static_cast<float>(tr.n - 1);
        }
        *out_n = n_stages;
    }
    for (cudaEvent_t e : tr.ev) cudaEventDestroy(e);
    return rc;
}

int afsim_selftest_math(AfsimHandle* h, uint64_t n, uint64_t out_mismatches[6]) {
    if (!h || !out_mismatches) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    AF_CUDA(h, cudaSetDevice(h->device));
    DeviceBuffers mem;
    unsigned long long* d = nullptr;
    AF_CUDA(h, mem.alloc(&d, 6));
    AF_CUDA(h, cudaMemsetAsync(d, 0, 6 * sizeof(unsigned long long), h->stream));
    AF_CUDA(h, launch_selftest_math(n, d, h->stream));
    unsigned long long host[6] = {0, 0, 0, 0, 0, 0};
    AF_CUDA(h, cudaMemcpyAsync(host, d, sizeof host, cudaMemcpyDeviceToHost, h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 6; ++k) out_mismatches[k] = host[k];
    return AFSIM_OK;
}

int afsim_measure_issue_peak(AfsimHandle* h, int kind, double* out) {
    if (!h || !out || (kind != 0 && kind != 1)) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    AF_CUDA(h, cudaSetDevice(h->device));
    DeviceBuffers mem;
    double* sink = nullptr;
    AF_CUDA(h, mem.alloc(&sink, 1));
    cudaDeviceProp prop;
    AF_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
    const int blocks = prop.multiProcessorCount * 8, iters = 20000;
    cudaEvent_t e0, e1;
    AF_CUDA(h, cudaEventCreate(&e0));
    AF_CUDA(h, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, h->stream);
        launch_issue_peak(kind, iters, blocks, sink, h->stream);
        cudaEventRecord(e1, h->stream);
        cudaStreamSynchronize(h->stream);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double instr = static_cast<double>(blocks) * 256.0 * iters * 8.0 * (kind == 0 ? 2.0 : 1.0);
        if (rep > 0 && ms > 0.0f) best = std::max(best, instr / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return cuda_fail(h, err, "afsim_measure_issue_peak");
    *out = best;
    return AFSIM_OK;
}

int afsim_chain_sweep(AfsimHandle* h, const float* const* passages, const size_t* passage_len, size_t n_passages,
                      double sample_rate, const AfCandidate* candidates, size_t n_candidates, const uint32_t* pair_passage,
                      const uint32_t* pair_candidate, size_t n_pairs, AfChainMetrics* out_metrics, float* const* out_audio) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    bool want_audio = false;
    if (out_audio)
        for (size_t i = 0; i < n_pairs; ++i) want_audio = want_audio || out_audio[i] != nullptr;
    AfsimSweep* sweep = nullptr;
    int rc = afsim_sweep_prepare(h, passages, passage_len, n_passages, sample_rate, candidates, n_candidates, pair_passage,
                                 pair_candidate, n_pairs, want_audio ? 1 : 0, &sweep);
    if (rc != AFSIM_OK) return rc;
    rc = afsim_sweep_launch(h, sweep);
    if (rc == AFSIM_OK) rc = afsim_sweep_collect(h, sweep, out_metrics);
    if (rc == AFSIM_OK && want_audio)
        for (size_t i = 0; i < n_pairs && rc == AFSIM_OK; ++i)
            if (out_audio[i]) rc = afsim_sweep_collect_audio(h, sweep, i, out_audio[i], sweep->pair_len[i]);
    const std::string keep = h->error;
    afsim_sweep_release(h, sweep);
    h->error = keep;
    return rc;
}

// ---- single-stream entry points ----------------------------------------------------------------------------------

namespace {
struct SweepReleaser {  // every exit path of a call that owns a sweep goes through afsim_sweep_release (stream syncs first)
    void operator()(AfsimSweep* s) const {
        if (!s) return;
        const std::string keep = s->owner ? s->owner->error : std::string();
        AfsimHandle* owner = s->owner;
        afsim_sweep_release(owner, s);
        if (owner) owner->error = keep;
    }
};
typedef std::unique_ptr<AfsimSweep, SweepReleaser> SweepGuard;
}  // namespace

int afsim_chain_render(AfsimHandle* h, const float* audio, size_t n, double sample_rate, const AfBand bands[AFSIM_NUM_BANDS],
                       const AfChainSettings* settings, AfChainMetrics* out_metrics, float* out_audio) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    if (!bands || !settings || !out_metrics || (!audio && n)) return set_error(h, AFSIM_INVALID_ARGUMENT, "null argument");
    const auto started = std::chrono::steady_clock::now();
    AfCandidate cand;
    std::memcpy(cand.bands, bands, sizeof cand.bands);
    cand.settings = *settings;
    const float* passages[1] = {audio};
    const size_t lens[1] = {n};
    float* outs[1] = {out_audio};
    const int rc = afsim_chain_sweep(h, passages, lens, 1, sample_rate, &cand, 1, nullptr, nullptr, 1, out_metrics,
                                     out_audio ? outs : nullptr);
    if (rc == AFSIM_OK)  // python_api.rs:387,694-697: wall time of the whole call
        out_metrics->candidate_runtime_ms =
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - started).count();
    return rc;
}

int afsim_eq_response(AfsimHandle* h, const double* frequencies_hz, size_t n_freqs, const AfBand* bands, size_t n_sets,
                      int typed, double sample_rate, double* out_db) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    if ((!frequencies_hz && n_freqs) || (!bands && n_sets) || (!out_db && n_freqs && n_sets))
        return set_error(h, AFSIM_INVALID_ARGUMENT, "null argument");
    std::string msg;
    for (size_t s = 0; s < n_sets; ++s) {
        const int rc = typed ? validate_typed_bands(bands + s * AFSIM_NUM_BANDS, sample_rate, &msg)
                             : validate_legacy_response_bands(bands + s * AFSIM_NUM_BANDS, sample_rate, &msg);
        if (rc != AFSIM_OK) return set_error(h, rc, msg);
    }
    if (n_sets == 0 && (!std::isfinite(sample_rate) || sample_rate <= 0.0))
        return set_error(h, AFSIM_INVALID_ARGUMENT, "sample_rate must be finite and positive");
    {
        const int rc = validate_response_frequencies(frequencies_hz, n_freqs, sample_rate, &msg);
        if (rc != AFSIM_OK) return set_error(h, rc, msg);
    }
    if (n_freqs == 0 || n_sets == 0) return AFSIM_OK;
    if (n_freqs * n_sets > 0x7fffffffu) return set_error(h, AFSIM_INVALID_ARGUMENT, "response grid too large");
    AF_CUDA(h, cudaSetDevice(h->device));
    std::vector<double> coeffs(n_sets * kMaxSections * 5);
    std::vector<int> sections(n_sets * AFSIM_NUM_BANDS);
    for (size_t s = 0; s < n_sets; ++s)
        plan_eq_sections(bands + s * AFSIM_NUM_BANDS, typed != 0, sample_rate,
                         reinterpret_cast<double(*)[5]>(coeffs.data() + s * kMaxSections * 5), sections.data() + s * AFSIM_NUM_BANDS);
    DeviceBuffers mem;
    mem.pool = &h->pool;
    double *d_coeffs = nullptr, *d_freqs = nullptr, *d_out = nullptr;
    int* d_sections = nullptr;
    AF_CUDA(h, mem.alloc(&d_coeffs, coeffs.size()));
    AF_CUDA(h, mem.alloc(&d_sections, sections.size()));
    AF_CUDA(h, mem.alloc(&d_freqs, n_freqs));
    AF_CUDA(h, mem.alloc(&d_out, n_freqs * n_sets));
    AF_CUDA(h, cudaMemcpyAsync(d_coeffs, coeffs.data(), coeffs.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    AF_CUDA(h, cudaMemcpyAsync(d_sections, sections.data(), sections.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    AF_CUDA(h, cudaMemcpyAsync(d_freqs, frequencies_hz, n_freqs * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    AF_CUDA(h, launch_eq_response(d_coeffs, d_sections, d_freqs, static_cast<int>(n_freqs), static_cast<int>(n_sets), sample_rate,
                                  d_out, h->stream));
    AF_CUDA(h, cudaMemcpyAsync(out_db, d_out, n_freqs * n_sets * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    return AFSIM_OK;
}

// max of the configured response over 512 log-spaced points 20 Hz .. 20 kHz (lib.rs:252-262)
static int eq_render_max_response(AfsimHandle* h, const AfBand bands[AFSIM_NUM_BANDS], double sample_rate, double* out_max) {
    std::vector<double> freqs(512), resp(512);
    for (int i = 0; i < 512; ++i) freqs[i] = 20.0 * std::pow(20000.0 / 20.0, static_cast<double>(i) / 511.0);
    std::vector<double> coeffs(kMaxSections * 5);
    std::vector<int> sections(AFSIM_NUM_BANDS);
    plan_eq_sections(bands, true, sample_rate, reinterpret_cast<double(*)[5]>(coeffs.data()), sections.data());
    DeviceBuffers mem;
    mem.pool = &h->pool;
    double *d_coeffs = nullptr, *d_freqs = nullptr, *d_out = nullptr;
    int* d_sections = nullptr;
    AF_CUDA(h, mem.alloc(&d_coeffs, coeffs.size()));
    AF_CUDA(h, mem.alloc(&d_sections, sections.size()));
    AF_CUDA(h, mem.alloc(&d_freqs, freqs.size()));
    AF_CUDA(h, mem.alloc(&d_out, resp.size()));
    AF_CUDA(h, cudaMemcpyAsync(d_coeffs, coeffs.data(), coeffs.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    AF_CUDA(h, cudaMemcpyAsync(d_sections, sections.data(), sections.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    AF_CUDA(h, cudaMemcpyAsync(d_freqs, freqs.data(), freqs.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    AF_CUDA(h, launch_eq_response(d_coeffs, d_sections, d_freqs, 512, 1, sample_rate, d_out, h->stream));
    AF_CUDA(h, cudaMemcpyAsync(resp.data(), d_out, resp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    double max_resp = -std::numeric_limits<double>::infinity();
    for (double v : resp) max_resp = std::fmax(max_resp, v);
    *out_max = max_resp;
    return AFSIM_OK;
}

// One long passage: the time-parallel cascade (afsim_eqscan.h) instead of a batch of one stream.
static int eq_render_scan(AfsimHandle* h, const float* audio, size_t n, double sample_rate, const AfBand bands[AFSIM_NUM_BANDS],
                          const CandidatePlan& plan, AfEqRenderStats* out_stats, float* out_audio) {
    AF_CUDA(h, cudaSetDevice(h->device));
    int log2_len = 6;  // 64-sample segments; longer ones once the passage has more than 32768 of them
    while ((n >> log2_len) > 32768 && log2_len < 20) ++log2_len;
    const size_t len = size_t(1) << log2_len, n_seg = (n + len - 1) / len;
    DeviceBuffers mem;
    mem.pool = &h->pool;
    float *d_in = nullptr, *d_out = nullptr, *d_xt = nullptr;
    double* d_state = nullptr;
    EqScanStats *d_partial = nullptr, *d_stats = nullptr;
    AF_CUDA(h, mem.alloc(&d_in, n));
    AF_CUDA(h, mem.alloc(&d_out, n));
    AF_CUDA(h, mem.alloc(&d_xt, n_seg * len));
    AF_CUDA(h, mem.alloc(&d_state, 4 * n_seg));
    AF_CUDA(h, mem.alloc(&d_partial, eqscan_stats_partials(n)));
    AF_CUDA(h, mem.alloc(&d_stats, 2));
    cudaEvent_t e0, e1;
    AF_CUDA(h, cudaEventCreate(&e0));
    AF_CUDA(h, cudaEventCreate(&e1));
    struct EventGuard {
        cudaEvent_t a, b;
        ~EventGuard() {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    } guard{e0, e1};
    AF_CUDA(h, cudaMemcpyAsync(d_in, audio, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    AF_CUDA(h, cudaEventRecord(e0, h->stream));
    int launches = 0;
    AF_CUDA(h, launch_plain_stats(d_in, n, d_partial, d_stats, h->stream));
    AF_CUDA(h, launch_eqscan(d_in, d_out, n, plan.params.eq, static_cast<int>(plan.params.n_sections), log2_len, d_xt, d_state,
                             &launches, h->stream));
    AF_CUDA(h, launch_plain_stats(d_out, n, d_partial, d_stats + 1, h->stream));
    AF_CUDA(h, cudaEventRecord(e1, h->stream));
    EqScanStats st[2];
    AF_CUDA(h, cudaMemcpyAsync(st, d_stats, sizeof st, cudaMemcpyDeviceToHost, h->stream));
    if (out_audio) AF_CUDA(h, cudaMemcpyAsync(out_audio, d_out, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.0f;
    AF_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    double max_resp = 0.0;
    const int rc = eq_render_max_response(h, bands, sample_rate, &max_resp);
    if (rc != AFSIM_OK) return rc;
    std::memset(out_stats, 0, sizeof *out_stats);
    const double divisor = static_cast<double>(std::max<size_t>(n, 1));
    out_stats->input_sample_peak = st[0].peak;
    out_stats->output_sample_peak = st[1].peak;
    out_stats->input_true_peak = st[0].tp_peak;
    out_stats->output_true_peak = st[1].tp_peak;
    out_stats->input_rms = std::sqrt(st[0].sum / divisor);
    out_stats->output_rms = std::sqrt(st[1].sum / divisor);
    out_stats->max_response_db = max_resp;
    out_stats->runtime_ms = static_cast<double>(ms);
    out_stats->sample_count = n;
    out_stats->algorithmic_latency_samples = 0;
    out_stats->non_finite_output = st[1].non_finite;
    return AFSIM_OK;
}

int afsim_eq_render(AfsimHandle* h, const float* audio, size_t n, double sample_rate, const AfBand bands[AFSIM_NUM_BANDS],
                    AfEqRenderStats* out_stats, float* out_audio) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    if (!bands || !out_stats || (!audio && n)) return set_error(h, AFSIM_INVALID_ARGUMENT, "null argument");
    CandidatePlan plan;
    std::string msg;
    int rc = plan_eq_only(bands, sample_rate, &plan, &msg);
    if (rc != AFSIM_OK) return set_error(h, rc, msg);
    for (size_t i = 0; i < n; ++i)  // lib.rs:225-229
        if (!std::isfinite(audio[i])) return set_error(h, AFSIM_INVALID_ARGUMENT, "audio must contain only finite samples");
    // Long single passages take the time-parallel cascade (within the render tolerance of the serial walk); short
    // ones stay on the bit-exact stage kernels.  AFSIM_EQ_SCAN_MIN overrides the threshold (samples).
    if (n >= static_cast<size_t>(env_int("AFSIM_EQ_SCAN_MIN", 65536)))
        return eq_render_scan(h, audio, n, sample_rate, bands, plan, out_stats, out_audio);
    const float* passages[1] = {audio};
    const size_t lens[1] = {n};
    PassageSource src;
    src.host = passages;
    AfsimSweep* sweep = nullptr;
    rc = build_sweep(h, src, lens, 1, sample_rate, nullptr, 1, &plan, nullptr, nullptr, 1, out_audio ? 1 : 0, &sweep);
    if (rc != AFSIM_OK) return rc;
    SweepGuard guard(sweep);
    rc = afsim_sweep_launch(h, sweep);
    if (rc != AFSIM_OK) return rc;
    StreamAccum acc;
    AF_CUDA(h, cudaMemcpyAsync(&acc, sweep->d_accum_first, sizeof acc, cudaMemcpyDeviceToHost, h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    float ms = 0.0f;
    AF_CUDA(h, cudaEventElapsedTime(&ms, sweep->ev_start, sweep->ev_stop));
    if (out_audio && n) {
        rc = afsim_sweep_collect_audio(h, sweep, 0, out_audio, n);
        if (rc != AFSIM_OK) return rc;
    }
    double max_resp = 0.0;
    rc = eq_render_max_response(h, bands, sample_rate, &max_resp);
    if (rc != AFSIM_OK) return rc;
    std::memset(out_stats, 0, sizeof *out_stats);
    const double divisor = static_cast<double>(std::max<size_t>(n, 1));
    out_stats->input_sample_peak = acc.peak_in;
    out_stats->output_sample_peak = acc.peak_out;
    out_stats->input_true_peak = acc.peak_in_tp;
    out_stats->output_true_peak = acc.peak_out_tp;
    out_stats->input_rms = std::sqrt(acc.sum_in / divisor);
    out_stats->output_rms = std::sqrt(acc.sum_out / divisor);
    out_stats->max_response_db = max_resp;
    out_stats->runtime_ms = static_cast<double>(ms);
    out_stats->sample_count = n;
    out_stats->algorithmic_latency_samples = 0;
    out_stats->non_finite_output = acc.non_finite;
    guard.reset();
    return AFSIM_OK;
}

// ---- simulate_auto_makeup_control (python_api.rs:118-276) ------------------------------------------------------------

void afsim_auto_makeup_settings_default(AfAutoMakeupSettings* s) {
    if (!s) return;
    std::memset(s, 0, sizeof *s);
    s->threshold_db = -24.0;
    s->ratio = 3.0;
    s->attack_ms = 10.0;
    s->release_ms = 180.0;
    s->makeup_gain_db = 0.0;
    s->target_lufs = -18.0;
    s->vad_reliability = 1.0;
    s->adaptive_release = 1;
    s->sidechain_highpass_enabled = 1;
}

int afsim_auto_makeup_sweep(AfsimHandle* h, const float* const* audio, const size_t* len, size_t n_streams, double sample_rate,
                            const double* const* vad, const double* noise_floor_db, const double* noise_reliability,
                            const AfAutoMakeupSettings* settings, float* const* out_traces, float* const* out_audio) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    if (n_streams == 0) return AFSIM_OK;
    if (!audio || !len || !noise_floor_db || !noise_reliability || !settings || !out_traces)
        return set_error(h, AFSIM_INVALID_ARGUMENT, "null argument");
    std::vector<CandidatePlan> plans(n_streams);
    std::vector<uint32_t> idx(n_streams);
    bool any_vad = false, want_audio = false;
    for (size_t i = 0; i < n_streams; ++i) {
        std::string msg;
        const bool has_vad = vad && vad[i];
        const int rc = plan_makeup_control(settings[i], sample_rate, noise_floor_db[i], noise_reliability[i], has_vad, &plans[i], &msg);
        if (rc != AFSIM_OK) return set_error(h, rc, msg);
        if (!(plans[i].structure & ST_AUTO_MAKEUP))  // Compressor::new found no loudness meter for this rate (:180-190)
            return set_error(h, AFSIM_UNSUPPORTED, "auto makeup needs a sample rate the loudness meter supports "
                                                   "(8000, 16000, 32000, 44100, 48000, 88200 or 96000 Hz)");
        if (!audio[i] && len[i]) return set_error(h, AFSIM_INVALID_ARGUMENT, "null audio");
        if (has_vad) {
            const size_t blocks = (len[i] + AFSIM_MAKEUP_CONTROL_BLOCK - 1) / AFSIM_MAKEUP_CONTROL_BLOCK;
            for (size_t b = 0; b < blocks; ++b)
                if (!std::isfinite(vad[i][b]) || !(vad[i][b] >= 0.0 && vad[i][b] <= 1.0))
                    return set_error(h, AFSIM_INVALID_ARGUMENT, "VAD probabilities must be finite and between 0 and 1");
            any_vad = true;
        }
        if (out_audio && out_audio[i]) want_audio = true;
        idx[i] = static_cast<uint32_t>(i);
    }
    PassageSource src;
    src.host = audio;
    SweepOptions opt;
    opt.block_samples = AFSIM_MAKEUP_CONTROL_BLOCK;
    opt.vad = any_vad ? vad : nullptr;
    opt.makeup_rows = true;
    AfsimSweep* sweep = nullptr;
    int rc = build_sweep(h, src, len, n_streams, sample_rate, nullptr, n_streams, plans.data(), idx.data(), idx.data(), n_streams,
                         want_audio ? 1 : 0, &sweep, opt);
    if (rc != AFSIM_OK) return rc;
    SweepGuard guard(sweep);
    rc = afsim_sweep_launch(h, sweep);
    if (rc != AFSIM_OK) return rc;
    std::vector<float> rows, mk_rows;
    for (auto& bp : sweep->batches) {
        const BatchArgs& a = bp->args;
        const size_t n_rows = static_cast<size_t>(a.n_rows), sp = static_cast<size_t>(a.stride);
        if (n_rows == 0) continue;
        rows.resize(4 * n_rows * sp);
        mk_rows.resize(3 * n_rows * sp);
        AF_CUDA(h, cudaMemcpyAsync(rows.data(), a.rows, rows.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        AF_CUDA(h, cudaMemcpyAsync(mk_rows.data(), a.mk_rows, mk_rows.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        AF_CUDA(h, cudaStreamSynchronize(h->stream));
        for (int s = 0; s < a.n_streams; ++s) {
            float* out = out_traces[bp->members[s]];
            if (!out) continue;
            for (size_t r = 0; r < n_rows; ++r) {
                out[0 * n_rows + r] = mk_rows[(0 * n_rows + r) * sp + s];  // makeup_gain_db
                out[1 * n_rows + r] = mk_rows[(1 * n_rows + r) * sp + s];  // activity
                out[2 * n_rows + r] = mk_rows[(2 * n_rows + r) * sp + s];  // reliability
                out[3 * n_rows + r] = rows[(2 * n_rows + r) * sp + s];     // gain_reduction_db at the block end
                out[4 * n_rows + r] = rows[(0 * n_rows + r) * sp + s];     // input_rms_db
                out[5 * n_rows + r] = rows[(1 * n_rows + r) * sp + s];     // output_rms_db
            }
        }
    }
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (want_audio)
        for (size_t i = 0; i < n_streams; ++i)
            if (out_audio[i] && len[i]) {
                rc = afsim_sweep_collect_audio(h, sweep, i, out_audio[i], len[i]);
                if (rc != AFSIM_OK) return rc;
            }
    return AFSIM_OK;
}

int afsim_auto_makeup_control(AfsimHandle* h, const float* audio, size_t n, double sample_rate, const double* vad_probabilities,
                              size_t n_vad, double noise_floor_db, double noise_reliability, const AfAutoMakeupSettings* settings,
                              float* out_traces, float* out_audio) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    if (!settings || (!audio && n) || (!vad_probabilities && n_vad) || (!out_traces && n))
        return set_error(h, AFSIM_INVALID_ARGUMENT, "null argument");
    // argument checks in the reference's order (python_api.rs:136-166)
    if (!std::isfinite(sample_rate) || sample_rate <= 0.0)
        return set_error(h, AFSIM_INVALID_ARGUMENT, "sample_rate must be positive and finite");
    if (!std::isfinite(noise_floor_db) || !std::isfinite(noise_reliability) || !(noise_reliability >= 0.0 && noise_reliability <= 1.0))
        return set_error(h, AFSIM_INVALID_ARGUMENT, "noise evidence must be finite and reliability must be between 0 and 1");
    for (size_t i = 0; i < n_vad; ++i)
        if (!std::isfinite(vad_probabilities[i]) || !(vad_probabilities[i] >= 0.0 && vad_probabilities[i] <= 1.0))
            return set_error(h, AFSIM_INVALID_ARGUMENT, "VAD probabilities must be finite and between 0 and 1");
    const size_t block_count = (n + AFSIM_MAKEUP_CONTROL_BLOCK - 1) / AFSIM_MAKEUP_CONTROL_BLOCK;
    if (n_vad != 0 && n_vad != block_count)
        return set_error(h, AFSIM_INVALID_ARGUMENT, "expected " + std::to_string(block_count) +
                                                        " VAD probabilities at the 10 ms control cadence, got " + std::to_string(n_vad));
    const float* audios[1] = {audio};
    const size_t lens[1] = {n};
    const double* vads[1] = {n_vad ? vad_probabilities : nullptr};
    float* traces[1] = {out_traces};
    float* outs[1] = {out_audio};
    return afsim_auto_makeup_sweep(h, audios, lens, 1, sample_rate, vads, &noise_floor_db, &noise_reliability, settings, traces,
                                   out_audio ? outs : nullptr);
}

}  // extern "C"

// ---- product resampler simulator (afsim_resample.cu) --------------------------------------------------------------

void afsim_resampler_spec_default(AfResamplerSpec* out) {
    if (!out) return;
    std::memset(out, 0, sizeof *out);
    out->chunk_size = 1024;
    out->sinc_len = 128;
    out->window = AF_WINDOW_BLACKMAN;
}

int afsim_product_resampler_shape(const AfResamplerSpec* spec, size_t n_in, AfResamplerShape* out_shape, char* err, size_t err_capacity) {
    std::string msg;
    int rc = AFSIM_INVALID_ARGUMENT;
    if (!spec || !out_shape) {
        msg = "null argument";
    } else {
        ResamplePlan plan;
        rc = plan_resampler(*spec, n_in, false, false, &plan, &msg);
        if (rc == AFSIM_OK) *out_shape = plan.shape;
    }
    if (err && err_capacity) {
        std::strncpy(err, msg.c_str(), err_capacity - 1);
        err[err_capacity - 1] = 0;
    }
    return rc;
}

int afsim_product_resampler_plan(const AfResamplerSpec* spec, size_t n_in, double* out_table, int64_t* out_base, int32_t* out_phase,
                                 double* out_frac) {
    if (!spec) return AFSIM_INVALID_ARGUMENT;
    ResamplePlan plan;
    std::string msg;
    const int rc = plan_resampler(*spec, n_in, true, out_table != nullptr, &plan, &msg);
    if (rc != AFSIM_OK) return rc;
    if (out_table) std::memcpy(out_table, plan.table.data(), plan.table.size() * sizeof(double));
    for (size_t i = 0; i < plan.frames.size(); ++i) {
        if (out_base) out_base[i] = plan.frames[i].base;
        if (out_phase) out_phase[i] = plan.frames[i].sub;
        if (out_frac) out_frac[i] = plan.frames[i].frac;
    }
    return AFSIM_OK;
}

namespace {
int resampler_on_device(AfsimHandle* h, const AfResamplerSpec& spec, size_t n_in) {  // plan + table + frames resident
    if (h->resample.matches(spec, n_in)) return AFSIM_OK;
    std::string msg;
    ResamplePlan plan;
    const int rc = plan_resampler(spec, n_in, true, true, &plan, &msg);
    if (rc != AFSIM_OK) return set_error(h, rc, msg);
    h->resample.drop();
    AF_CUDA(h, cudaMalloc(&h->resample.d_table, plan.table.size() * sizeof(double)));
    AF_CUDA(h, cudaMalloc(&h->resample.d_frames, std::max<size_t>(plan.frames.size(), 1) * sizeof(ResampleFrame)));
    AF_CUDA(h, cudaMemcpyAsync(h->resample.d_table, plan.table.data(), plan.table.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    AF_CUDA(h, cudaMemcpyAsync(h->resample.d_frames, plan.frames.data(), plan.frames.size() * sizeof(ResampleFrame), cudaMemcpyHostToDevice,
                               h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));  // the host vectors go away below
    plan.frames.clear();
    plan.frames.shrink_to_fit();
    plan.table.clear();
    plan.table.shrink_to_fit();
    h->resample.plan = std::move(plan);
    return AFSIM_OK;
}
}  // namespace

int afsim_product_resampler_device(AfsimHandle* h, const AfResamplerSpec* spec, const double* d_in, size_t in_stride, size_t n_streams,
                                   size_t n_in, double* d_out, size_t out_stride, float* out_ms) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    if (!spec || (!d_in && n_in && n_streams) || (!d_out && n_streams)) return set_error(h, AFSIM_INVALID_ARGUMENT, "null argument");
    if (n_streams > 0x7fffffffu) return set_error(h, AFSIM_INVALID_ARGUMENT, "too many streams");
    AF_CUDA(h, cudaSetDevice(h->device));
    const int rc = resampler_on_device(h, *spec, n_in);
    if (rc != AFSIM_OK) return rc;
    const ResamplePlan& plan = h->resample.plan;
    if (in_stride < n_in || out_stride < plan.shape.frames) return set_error(h, AFSIM_INVALID_ARGUMENT, "stride shorter than the signal");
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (out_ms) {
        AF_CUDA(h, cudaEventCreate(&e0));
        AF_CUDA(h, cudaEventCreate(&e1));
        AF_CUDA(h, cudaEventRecord(e0, h->stream));
    }
    AF_CUDA(h, launch_resample(d_in, in_stride, n_in, d_out, out_stride, plan.shape.frames, static_cast<int>(n_streams), h->resample.d_frames,
                               h->resample.d_table, static_cast<int>(spec->sinc_len), plan.max_span, h->stream));
    if (out_ms) AF_CUDA(h, cudaEventRecord(e1, h->stream));
    AF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (out_ms) {
        AF_CUDA(h, cudaEventElapsedTime(out_ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    return AFSIM_OK;
}

int afsim_product_resampler(AfsimHandle* h, const AfResamplerSpec* spec, const double* const* samples, size_t n_streams, size_t n_in,
                            double* const* out, AfResamplerShape* out_shape) {
    if (!h) return AFSIM_INVALID_ARGUMENT;
    AF_LOCK(h);
    h->error.clear();
    if (!spec || (!samples && n_streams) || (!out && n_streams)) return set_error(h, AFSIM_INVALID_ARGUMENT, "null argument");
    if (n_streams > 0x7fffffffu) return set_error(h, AFSIM_INVALID_ARGUMENT, "too many streams");
    {   // validation in the reference's order (resampling.rs:187-221): rates, chunk, sinc_len, window, then the samples
        std::string msg;
        ResamplePlan probe;
        const int rc = plan_resampler(*spec, n_in, false, false, &probe, &msg);
        if (rc == AFSIM_INVALID_ARGUMENT && msg.rfind("resampler flush", 0) != 0) return set_error(h, rc, msg);
        for (size_t s = 0; s < n_streams; ++s) {
            if (!samples[s] && n_in) return set_error(h, AFSIM_INVALID_ARGUMENT, "null argument");
            for (size_t i = 0; i < n_in; ++i)
                if (!std::isfinite(samples[s][i])) return set_error(h, AFSIM_INVALID_ARGUMENT, "samples must be finite");
        }
        if (rc != AFSIM_OK) return set_error(h, rc, msg);
        if (out_shape) *out_shape = probe.shape;
    }
    if (n_streams == 0) return AFSIM_OK;
    AF_CUDA(h, cudaSetDevice(h->device));
    const int rc = resampler_on_device(h, *spec, n_in);
    if (rc != AFSIM_OK) return rc;
    const ResamplePlan& plan = h->resample.plan;
    const size_t frames = plan.shape.frames;
    const size_t in_stride = (n_in + 31) / 32 * 32 + 32, out_stride = (frames + 31) / 32 * 32;
    DeviceBuffers mem;
    mem.pool = &h->pool;
    struct Quiesce {  // whatever way this call ends, nothing is in flight when `mem` hands its buffers back to the pool
        cudaStream_t st;
        ~Quiesce() { cudaStreamSynchronize(st); }
    } quiesce{h->stream};
    // Signals go through the device in groups; every group is staged through one of two PINNED host areas (inputs and outputs
    // side by side), filled and emptied by a few host threads while the other area's group is on the bus / the GPU: pageable
    // cudaMemcpy moved the 1.4 GB of a 32 x 60 s call at ~4 GB/s, which was 98 % of the call.
    const size_t host_per_stream = (n_in + frames) * sizeof(double);
    size_t group = std::max<size_t>(1, std::min(n_streams, (size_t(384) << 20) / std::max<size_t>(host_per_stream, 1)));
    if (group >= 8) group = group / 8 * 8;  // whole groups of eight for the kernel
    ResampleCache& rc_ = h->resample;
    if (rc_.pinned_bytes < group * host_per_stream || !rc_.done[0]) {
        rc_.drop_staging();
        const size_t want = std::max<size_t>(group * host_per_stream, size_t(1) << 20);
        for (int i = 0; i < 2; ++i) {
            AF_CUDA(h, cudaHostAlloc(&rc_.pinned[i], want, cudaHostAllocDefault));
            AF_CUDA(h, cudaEventCreateWithFlags(&rc_.done[i], cudaEventDisableTiming));
        }
        rc_.pinned_bytes = want;
    }
    double *d_in[2] = {nullptr, nullptr}, *d_out[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i) {
        AF_CUDA(h, mem.alloc(&d_in[i], group * in_stride));
        AF_CUDA(h, mem.alloc(&d_out[i], group * std::max<size_t>(out_stride, 1)));
    }
    const unsigned n_threads = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    auto parallel_rows = [&](size_t rows, size_t row_bytes, auto&& row_copy) {  // row_copy(row, offset, bytes)
        if (rows * row_bytes < (size_t(4) << 20) || n_threads == 1) {
            for (size_t r = 0; r < rows; ++r) row_copy(r, size_t(0), row_bytes);
            return;
        }
        std::vector<std::thread> pool;
        const size_t total = rows * row_bytes, per = (total / n_threads + 4095) / 4096 * 4096;
        for (unsigned t = 0; t < n_threads; ++t) {
            const size_t lo = t * per, hi = std::min(total, lo + per);
            if (lo >= hi) break;
            pool.emplace_back([&, lo, hi] {
                for (size_t pos = lo; pos < hi;) {
                    const size_t r = pos / row_bytes, off = pos % row_bytes, len = std::min(row_bytes - off, hi - pos);
                    row_copy(r, off, len);
                    pos += len;
                }
            });
        }
        for (auto& th : pool) th.join();
    };
    const size_t n_groups = (n_streams + group - 1) / group;
    auto first_of = [&](size_t g) { return g * group; };
    auto count_of = [&](size_t g) { return std::min(group, n_streams - g * group); };
    auto drain = [&](size_t g) -> int {  // results of group g: pinned -> the caller's buffers
        const int x = static_cast<int>(g & 1);
        AF_CUDA(h, cudaEventSynchronize(rc_.done[x]));
        const char* src = static_cast<const char*>(rc_.pinned[x]) + group * n_in * sizeof(double);
        const size_t s0 = first_of(g);
        parallel_rows(count_of(g), frames * sizeof(double), [&](size_t r, size_t off, size_t len) {
            std::memcpy(reinterpret_cast<char*>(out[s0 + r]) + off, src + r * frames * sizeof(double) + off, len);
        });
        return AFSIM_OK;
    };
    for (size_t s = 0; s < n_streams; ++s)
        if (!out[s]) return set_error(h, AFSIM_INVALID_ARGUMENT, "null argument");
    for (size_t g = 0; g < n_groups; ++g) {
        const int x = static_cast<int>(g & 1);
        const size_t s0 = first_of(g), ns = count_of(g);
        if (g >= 2) {
            const int rc2 = drain(g - 2);
            if (rc2 != AFSIM_OK) return rc2;
        }
        char* pin_in = static_cast<char*>(rc_.pinned[x]);
        char* pin_out = pin_in + group * n_in * sizeof(double);
        if (n_in) {
            parallel_rows(ns, n_in * sizeof(double), [&](size_t r, size_t off, size_t len) {
                std::memcpy(pin_in + r * n_in * sizeof(double) + off, reinterpret_cast<const char*>(samples[s0 + r]) + off, len);
            });
            AF_CUDA(h, cudaMemcpy2DAsync(d_in[x], in_stride * sizeof(double), pin_in, n_in * sizeof(double), n_in * sizeof(double), ns,
                                         cudaMemcpyHostToDevice, h->stream));
        }
        AF_CUDA(h, launch_resample(d_in[x], in_stride, n_in, d_out[x], out_stride, frames, static_cast<int>(ns), h->resample.d_frames,
                                   h->resample.d_table, static_cast<int>(spec->sinc_len), plan.max_span, h->stream));
        if (frames)
            AF_CUDA(h, cudaMemcpy2DAsync(pin_out, frames * sizeof(double), d_out[x], out_stride * sizeof(double), frames * sizeof(double), ns,
                                         cudaMemcpyDeviceToHost, h->stream));
        AF_CUDA(h, cudaEventRecord(rc_.done[x], h->stream));
    }
    for (size_t g = n_groups >= 2 ? n_groups - 2 : 0; g < n_groups; ++g) {
        const int rc2 = drain(g);
        if (rc2 != AFSIM_OK) return rc2;
    }
    return AFSIM_OK;
}
