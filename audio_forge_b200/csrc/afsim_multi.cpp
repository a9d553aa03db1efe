// afsim_multi.cpp -- multi-GPU sweeps behind the C ABI (include/afsim.h: afsim_multi_*).
//
// The Rust host binds ONE handle for the whole box (the shape of the reference's runtime-loaded plugin,
// dsp/deepfilter_ffi.rs:335-386): afsim_multi_create(device_mask) opens one AfsimHandle per selected GPU and one NCCL
// communicator over them (ncclCommInitAll: a single process drives all devices).  afsim_multi_chain_sweep partitions
// the candidate x passage streams over the GPUs by cost (longest-processing-time on (fixed chain cost + EQ sections)
// x samples -- the rule of audio_forge_b200/sharding.py), renders every shard on its own host thread through the
// ordinary single-GPU sweep, and all-gathers the per-stream AfChainMetrics structs straight from the sweeps' device
// tables with ncclAllGather (grouped, device to device over NVLink); the table leaves the GPUs once, from the first
// device, and is permuted into the caller's pair order.  Streams are independent end to end, so this gather is the
// only exchange on the path.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): a single-GPU user of libafsim.so needs no NCCL at all, and
// inside a Python process that already loaded torch's NCCL the same library instance is reused.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/afsim.h"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string* why) {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) {
            *why = std::string("NCCL is not available: ") + dlerror();
            return false;
        }
        auto sym = [&](const char* name) { return dlsym(lib, name); };
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(sym("ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(sym("ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(sym("ncclGroupEnd"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !AllGather || !GetErrorString) {
            *why = "libnccl.so.2 lacks a required symbol";
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
thread_local std::string g_multi_create_error;

}  // namespace

struct AfsimMulti {
    std::vector<int> devices;
    std::vector<AfsimHandle*> handles;
    std::vector<cudaStream_t> streams;   // one per device: the handle launches on it, NCCL gathers on it
    std::vector<ncclComm_t> comms;       // empty for a single device
    std::vector<unsigned char*> gathered;  // per device: [n_dev * cap] bytes
    std::vector<unsigned char*> padded;    // per device: [cap] bytes (the device's shard, padded to the largest shard)
    size_t cap_bytes = 0;
    std::string error;
    std::mutex mu;
};

namespace {

int fail(AfsimMulti* m, int status, const std::string& msg) {
    if (m) m->error = msg;
    return status;
}

// EQ sections a candidate renders (the cost model of sharding.stream_costs)
double candidate_sections(const AfCandidate& c) {
    int n = 0;
    for (int b = 0; b < AFSIM_NUM_BANDS; ++b) {
        const AfBand& band = c.bands[b];
        if (!c.settings.use_typed_bands)
            n += 1;
        else if (band.enabled)
            n += (band.filter_type == 4 || band.filter_type == 5) ? band.slope_db_per_octave / 12 : 1;
    }
    return static_cast<double>(n);
}

}  // namespace

extern "C" {

const char* afsim_multi_create_error(void) { return g_multi_create_error.c_str(); }

int afsim_multi_create(uint32_t device_mask, AfsimMulti** out) {
    g_multi_create_error.clear();
    if (!out) {
        g_multi_create_error = "out is null";
        return AFSIM_INVALID_ARGUMENT;
    }
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        g_multi_create_error = "no CUDA device (libafsim has no CPU path)";
        return AFSIM_CUDA_ERROR;
    }
    auto m = new AfsimMulti();
    for (int d = 0; d < count && d < 32; ++d)
        if (device_mask & (1u << d)) m->devices.push_back(d);
    if (m->devices.empty() || (count < 32 && (device_mask >> count) != 0)) {
        g_multi_create_error = "device_mask selects no device or a device that does not exist";
        delete m;
        return AFSIM_INVALID_ARGUMENT;
    }
    const int n = static_cast<int>(m->devices.size());
    for (int i = 0; i < n; ++i) {
        cudaStream_t st = nullptr;
        AfsimHandle* h = nullptr;
        int rc = AFSIM_CUDA_ERROR;
        if (cudaSetDevice(m->devices[i]) == cudaSuccess && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess)
            rc = afsim_create(m->devices[i], st, &h);
        else
            g_multi_create_error = "cudaStreamCreate failed";
        if (rc != AFSIM_OK) {
            if (g_multi_create_error.empty()) g_multi_create_error = afsim_create_error();
            if (st) cudaStreamDestroy(st);
            afsim_multi_destroy(m);
            return rc;
        }
        m->streams.push_back(st);
        m->handles.push_back(h);
    }
    if (n > 1) {
        std::lock_guard<std::mutex> lock(g_nccl_mu);
        std::string why;
        if (!g_nccl.load(&why)) {
            g_multi_create_error = why;
            afsim_multi_destroy(m);
            return AFSIM_UNSUPPORTED;
        }
        m->comms.assign(n, nullptr);
        const ncclResult_t r = g_nccl.CommInitAll(m->comms.data(), n, m->devices.data());
        if (r != ncclSuccess) {
            g_multi_create_error = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r);
            m->comms.clear();
            afsim_multi_destroy(m);
            return AFSIM_CUDA_ERROR;
        }
    }
    *out = m;
    return AFSIM_OK;
}

void afsim_multi_destroy(AfsimMulti* m) {
    if (!m) return;
    for (size_t i = 0; i < m->handles.size(); ++i) {
        cudaSetDevice(m->devices[i]);
        if (i < m->gathered.size() && m->gathered[i]) cudaFree(m->gathered[i]);
        if (i < m->padded.size() && m->padded[i]) cudaFree(m->padded[i]);
    }
    for (ncclComm_t c : m->comms)
        if (c) g_nccl.CommDestroy(c);
    for (size_t i = 0; i < m->handles.size(); ++i) {
        afsim_destroy(m->handles[i]);
        cudaSetDevice(m->devices[i]);
        cudaStreamDestroy(m->streams[i]);
    }
    delete m;
}

int afsim_multi_device_count(const AfsimMulti* m) { return m ? static_cast<int>(m->devices.size()) : 0; }

const char* afsim_multi_last_error(const AfsimMulti* m) { return m ? m->error.c_str() : "null handle"; }

int afsim_multi_partition(const AfCandidate* candidates, size_t n_candidates, const size_t* passage_len, size_t n_passages,
                          const uint32_t* pair_passage, const uint32_t* pair_candidate, size_t n_pairs, int n_parts,
                          uint32_t* out_owner) {
    if (!out_owner || n_parts < 1 || (n_pairs && (!candidates || !passage_len))) return AFSIM_INVALID_ARGUMENT;
    if ((pair_passage == nullptr) != (pair_candidate == nullptr)) return AFSIM_INVALID_ARGUMENT;
    std::vector<double> sections(n_candidates);
    for (size_t c = 0; c < n_candidates; ++c) sections[c] = candidate_sections(candidates[c]);
    std::vector<double> cost(n_pairs);
    for (size_t i = 0; i < n_pairs; ++i) {
        const size_t p = pair_passage ? pair_passage[i] : i % n_passages;
        const size_t c = pair_candidate ? pair_candidate[i] : i / n_passages;
        if (p >= n_passages || c >= n_candidates) return AFSIM_INVALID_ARGUMENT;
        cost[i] = (40.0 + sections[c]) * static_cast<double>(passage_len[p]);
    }
    // Units of the split: the streams of one candidate, in caller order, in pieces of at most ceil(n / (16 parts)) -- a rank
    // that holds all passages of its candidates reads 1 / parts of the candidate constants (sharding.shard_units).
    const size_t cap = std::max<size_t>(1, (n_pairs + 16 * static_cast<size_t>(n_parts) - 1) / (16 * static_cast<size_t>(n_parts)));
    std::vector<uint32_t> unit_of(n_pairs);
    std::vector<int64_t> open_unit(n_candidates, -1);
    std::vector<double> unit_cost;
    std::vector<size_t> unit_size;
    for (size_t i = 0; i < n_pairs; ++i) {
        const size_t c = pair_candidate ? pair_candidate[i] : i / n_passages;
        int64_t u = open_unit[c];
        if (u < 0 || unit_size[u] >= cap) {
            u = static_cast<int64_t>(unit_cost.size());
            open_unit[c] = u;
            unit_cost.push_back(0.0);
            unit_size.push_back(0);
        }
        unit_of[i] = static_cast<uint32_t>(u);
        unit_cost[u] += cost[i];
        unit_size[u] += 1;
    }
    std::vector<uint32_t> order(unit_cost.size());
    for (size_t u = 0; u < order.size(); ++u) order[u] = static_cast<uint32_t>(u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return unit_cost[a] > unit_cost[b]; });
    typedef std::tuple<double, size_t, int> Load;  // (load, streams, part): least loaded, then fewest streams, then lowest index
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (int r = 0; r < n_parts; ++r) heap.emplace(0.0, size_t(0), r);
    std::vector<uint32_t> unit_owner(unit_cost.size());
    for (uint32_t u : order) {
        Load top = heap.top();
        heap.pop();
        unit_owner[u] = static_cast<uint32_t>(std::get<2>(top));
        heap.emplace(std::get<0>(top) + unit_cost[u], std::get<1>(top) + unit_size[u], std::get<2>(top));
    }
    for (size_t i = 0; i < n_pairs; ++i) out_owner[i] = unit_owner[unit_of[i]];
    return AFSIM_OK;
}

int afsim_multi_chain_sweep(AfsimMulti* m, const float* const* passages, const size_t* passage_len, size_t n_passages,
                            double sample_rate, const AfCandidate* candidates, size_t n_candidates,
                            const uint32_t* pair_passage, const uint32_t* pair_candidate, size_t n_pairs,
                            AfChainMetrics* out_metrics, float* out_device_ms) {
    if (!m) return AFSIM_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(m->mu);
    m->error.clear();
    if (n_pairs && !out_metrics) return fail(m, AFSIM_INVALID_ARGUMENT, "out_metrics is null");
    if (!pair_passage && n_pairs != n_passages * n_candidates)
        return fail(m, AFSIM_INVALID_ARGUMENT, "n_pairs must equal n_candidates * n_passages for a full cross product");
    const int n_dev = static_cast<int>(m->devices.size());
    if (n_pairs == 0) return AFSIM_OK;

    // ---- partition -------------------------------------------------------------------------------------------------
    std::vector<uint32_t> owner(n_pairs);
    int rc = afsim_multi_partition(candidates, n_candidates, passage_len, n_passages, pair_passage, pair_candidate, n_pairs, n_dev,
                                   owner.data());
    if (rc != AFSIM_OK) return fail(m, rc, "pair index out of range");
    std::vector<std::vector<uint32_t>> shard(n_dev), pp(n_dev), pc(n_dev);
    for (size_t i = 0; i < n_pairs; ++i) {
        const int d = static_cast<int>(owner[i]);
        shard[d].push_back(static_cast<uint32_t>(i));
        pp[d].push_back(pair_passage ? pair_passage[i] : static_cast<uint32_t>(i % n_passages));
        pc[d].push_back(pair_candidate ? pair_candidate[i] : static_cast<uint32_t>(i / n_passages));
    }
    size_t max_n = 0;
    for (int d = 0; d < n_dev; ++d) max_n = std::max(max_n, shard[d].size());
    const size_t cap = max_n * sizeof(AfChainMetrics);

    // ---- gather buffers (grown on demand, reused between calls) ------------------------------------------------------------
    if (n_dev > 1 && cap > m->cap_bytes) {
        m->gathered.resize(n_dev, nullptr);
        m->padded.resize(n_dev, nullptr);
        for (int d = 0; d < n_dev; ++d) {
            cudaSetDevice(m->devices[d]);
            if (m->gathered[d]) cudaFree(m->gathered[d]);
            if (m->padded[d]) cudaFree(m->padded[d]);
            m->gathered[d] = m->padded[d] = nullptr;
            if (cudaMalloc(&m->gathered[d], cap * n_dev) != cudaSuccess || cudaMalloc(&m->padded[d], cap) != cudaSuccess) {
                cudaGetLastError();
                m->cap_bytes = 0;
                return fail(m, AFSIM_OUT_OF_MEMORY, "gather buffers");
            }
        }
        m->cap_bytes = cap;
    }

    // ---- render: one host thread per device ----------------------------------------------------------------------------------
    std::vector<AfsimSweep*> sweeps(n_dev, nullptr);
    std::vector<int> status(n_dev, AFSIM_OK);
    std::vector<cudaEvent_t> ev0(n_dev), ev1(n_dev);
    auto render = [&](int d) {
        AfsimHandle* h = m->handles[d];
        cudaSetDevice(m->devices[d]);
        cudaEventCreate(&ev0[d]);
        cudaEventCreate(&ev1[d]);
        if (shard[d].empty()) return;
        status[d] = afsim_sweep_prepare(h, passages, passage_len, n_passages, sample_rate, candidates, n_candidates, pp[d].data(),
                                        pc[d].data(), shard[d].size(), 0, &sweeps[d]);
        if (status[d] != AFSIM_OK) return;
        cudaEventRecord(ev0[d], m->streams[d]);
        status[d] = afsim_sweep_launch(h, sweeps[d]);
        if (status[d] == AFSIM_OK && n_dev > 1) {
            // the shard's metric table, padded to the largest shard: NCCL all-gather wants equal counts
            cudaMemsetAsync(m->padded[d], 0, cap, m->streams[d]);
            cudaMemcpyAsync(m->padded[d], afsim_sweep_metrics_device_ptr(sweeps[d]), shard[d].size() * sizeof(AfChainMetrics),
                            cudaMemcpyDeviceToDevice, m->streams[d]);
        }
    };
    {
        std::vector<std::thread> workers;
        for (int d = 1; d < n_dev; ++d) workers.emplace_back(render, d);
        render(0);
        for (std::thread& t : workers) t.join();
    }
    auto cleanup = [&]() {
        for (int d = 0; d < n_dev; ++d) {
            cudaSetDevice(m->devices[d]);
            if (sweeps[d]) afsim_sweep_release(m->handles[d], sweeps[d]);
            cudaEventDestroy(ev0[d]);
            cudaEventDestroy(ev1[d]);
        }
    };
    for (int d = 0; d < n_dev; ++d)
        if (status[d] != AFSIM_OK) {
            const std::string msg = afsim_last_error(m->handles[d]);
            cleanup();
            return fail(m, status[d], "device " + std::to_string(m->devices[d]) + ": " + msg);
        }

    // ---- the one collective: all-gather of the metric structs, device to device -------------------------------------------------
    std::vector<AfChainMetrics> table;
    if (n_dev > 1) {
        ncclResult_t r = g_nccl.GroupStart();
        for (int d = 0; d < n_dev && r == ncclSuccess; ++d)
            r = g_nccl.AllGather(m->padded[d], m->gathered[d], cap, ncclUint8, m->comms[d], m->streams[d]);
        const ncclResult_t r2 = g_nccl.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) {
            cleanup();
            return fail(m, AFSIM_CUDA_ERROR, std::string("ncclAllGather: ") + g_nccl.GetErrorString(r));
        }
        for (int d = 0; d < n_dev; ++d) {
            cudaSetDevice(m->devices[d]);
            cudaEventRecord(ev1[d], m->streams[d]);
        }
        table.resize(max_n * n_dev);
        cudaSetDevice(m->devices[0]);
        cudaError_t err = cudaMemcpyAsync(table.data(), m->gathered[0], cap * n_dev, cudaMemcpyDeviceToHost, m->streams[0]);
        for (int d = 0; d < n_dev && err == cudaSuccess; ++d) {
            cudaSetDevice(m->devices[d]);
            err = cudaStreamSynchronize(m->streams[d]);
        }
        if (err != cudaSuccess) {
            cleanup();
            return fail(m, AFSIM_CUDA_ERROR, std::string("gather: ") + cudaGetErrorString(err));
        }
        for (int d = 0; d < n_dev; ++d)
            for (size_t k = 0; k < shard[d].size(); ++k) out_metrics[shard[d][k]] = table[static_cast<size_t>(d) * max_n + k];
    } else {
        table.resize(n_pairs);
        cudaEventRecord(ev1[0], m->streams[0]);
        rc = afsim_sweep_collect(m->handles[0], sweeps[0], table.data());
        if (rc != AFSIM_OK) {
            const std::string msg = afsim_last_error(m->handles[0]);
            cleanup();
            return fail(m, rc, msg);
        }
        for (size_t k = 0; k < n_pairs; ++k) out_metrics[shard[0][k]] = table[k];
    }
    if (n_dev > 1)  // asynchronous failures of a device's launch (afsim_sweep_collect reports them on the single-device path)
        for (int d = 0; d < n_dev; ++d) {
            if (!sweeps[d]) continue;
            rc = afsim_sweep_status(m->handles[d], sweeps[d]);
            if (rc != AFSIM_OK) {
                const std::string msg = afsim_last_error(m->handles[d]);
                cleanup();
                return fail(m, rc, "device " + std::to_string(m->devices[d]) + ": " + msg);
            }
        }
    float worst_ms = 0.0f;
    for (int d = 0; d < n_dev; ++d) {
        cudaSetDevice(m->devices[d]);
        float ms = 0.0f;
        if (!shard[d].empty() && cudaEventElapsedTime(&ms, ev0[d], ev1[d]) == cudaSuccess) worst_ms = std::max(worst_ms, ms);
    }
    if (out_device_ms) *out_device_ms = worst_ms;
    for (size_t i = 0; i < n_pairs; ++i) out_metrics[i].candidate_runtime_ms = static_cast<double>(worst_ms) / n_pairs;
    cleanup();
    return AFSIM_OK;
}

}  // extern "C"
