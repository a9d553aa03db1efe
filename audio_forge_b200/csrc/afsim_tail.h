// afsim_tail.h -- launcher of the fused tail kernel (afsim_tail.cu): sample limiter -> true-peak limiter ->
// true-peak detector + output statistics of one chunk in one SM-local, TMA-fed kernel.
#pragma once
#include <cuda_runtime.h>

#include "afsim_params.h"

namespace afsim {

struct ChunkArgs;

constexpr int kTailMapWarps = 4;
constexpr int kTailThreads = 32 * (4 + kTailMapWarps);  // LIM-R (+ TMA producer), TP-R, two LIM-M warps, four FIR warps
constexpr int kTailHistRows = 64;                      // parked between chunks: 32 rows of each of the two output rings

struct alignas(64) TailMap {  // a CUtensorMap (kept opaque so that callers need not include <cuda.h>)
    unsigned char bytes[128];
};

// lookaheads the kernel's shared-memory x ring can hold (<= 28 sub-tiles of 32 samples of lookback)
bool tail_supported(int lookahead);
// 2-D tensor map of a stream-minor f32 ring [ring_rows][stride]: box = 32 time rows x 32 streams
cudaError_t tail_make_map(const float* ring, int ring_rows, int stride, TailMap* out);
cudaError_t tail_configure();
// err_flag: device int, set to 1 if the kernel's pipeline watchdog ever fired (a bug; results are then invalid)
cudaError_t launch_tail(const BatchArgs& a, const ChunkArgs& ck, const TailMap& map, int* err_flag, cudaStream_t st);

}  // namespace afsim
