// afsim_kernels.h -- launchers of the stage kernels (afsim_kernels.cu), called by afsim_api.cu.
#pragma once
#include <cuda_runtime.h>

#include "afsim_params.h"

namespace afsim {

struct ChunkArgs;

constexpr int kFinalizeThreads = 128;
constexpr int kMapWarps = 4;  // warps per block of the map (M) kernels
constexpr int kRBlock = 32;  // threads per block of the serial (R) kernels: one warp, so few-stream batches reach every SM
constexpr size_t kFinalizeSmemLimit = 200 * 1024;
constexpr int kFirBlocksPerSm = 8;  // grid cap of the FIR / limiter-window maps in blocks per SM (0 = none); see launch_split
constexpr int kWarpPerStreamMax = 64;  // input cleanup: batches up to this many streams run one warp per stream (32x the warps: only where latency, not issue load, matters)

cudaError_t launch_expand_deesser(const BatchArgs& a, cudaStream_t st);
cudaError_t launch_input(const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st);
cudaError_t launch_input_fanout(const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st);
cudaError_t launch_eq(const BatchArgs& a, const ChunkArgs& ck, int first_section, int k, cudaStream_t st);
cudaError_t launch_compressor(const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st);
cudaError_t launch_limiter(const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st);
cudaError_t launch_output(const BatchArgs& a, const ChunkArgs& ck, bool limiter, cudaStream_t st);
cudaError_t launch_input_true_peak(const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st);
// split (R/M) path, afsim_split.h
enum SplitOp {
    SP_COMP_R1, SP_COMP_M2, SP_COMP_R3, SP_COMP_M4, SP_COMP_R5, SP_COMP_M6,
    SP_LIM_M, SP_LIM_R, SP_TP_FIR_IN, SP_TP_R, SP_TP_FIR_OUT,
    SP_DE_RA, SP_DE_MB, SP_DE_RC,
    SP_COMP_R7,  // auto makeup (after M6)
    SP_DE_MC2, SP_DE_RC3,  // de-esser: coefficient rebuild map, dynamic-EQ biquads (SP_DE_RC is R_c1)
    SP_DE_RC1A, SP_DE_MC1B, SP_DE_RC1C  // R_c1 cut three ways for few-stream batches (afsim_deesser.h)
};
cudaError_t launch_split(SplitOp op, const BatchArgs& a, const ChunkArgs& ck, cudaStream_t st);
cudaError_t configure_kernels();
cudaError_t launch_finalize(const BatchArgs& a, cudaStream_t st);
size_t finalize_workspace_bytes(int n_rows, int n_pad);
cudaError_t launch_eq_response(const double* coeffs, const int* n_sections, const double* freqs, int n_freqs,
                               int n_sets, double fs, double* out, cudaStream_t st);
// time-parallel EQ render of one long passage + plain-array statistics (afsim_eqscan.cu)
struct EqScanStats {  // layout of PlainStats
    double sum;
    float peak, tp_peak;
    unsigned non_finite, pad;
};
size_t eqscan_stats_partials(size_t n);
cudaError_t launch_plain_stats(const float* x, size_t n, void* partial, void* out, cudaStream_t st);
cudaError_t launch_eqscan(const float* d_in, float* d_out, size_t n, const double (*coeffs)[5], int n_sections, int log2_len,
                          float* xt, double* seg_state, int* launches, cudaStream_t st);
cudaError_t launch_selftest_math(unsigned long long n, unsigned long long* counts, cudaStream_t st);
cudaError_t launch_issue_peak(int kind, int iters, int blocks, double* sink, cudaStream_t st);
cudaError_t launch_synth(float* out, size_t n_per, int n_passages, int kind, double fs, cudaStream_t st);

}  // namespace afsim
