// Product resampler simulator (reference: rust-core/src/audio/processor/resampling.rs:140-262, `simulate_product_resampler`
// over rubato 0.14 `SincFixedIn<f64>`, cubic interpolation between 256 windowed-sinc phases).
// Host planner (phase table, frame positions: the walk of the reference's block loop) + the device kernel's launcher.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/afsim.h"

namespace afsim {

constexpr int kResamplePhases = 256;       // oversampling_factor (resampling.rs:152)
constexpr int kResampleFramesPerBlock = 128;  // frames one CTA renders (4 warps x 32)

struct ResampleFrame {  // one output frame: which input window, which phase, where between two phases
    double frac;        // position between phase `sub` and `sub + 1` (interp_cubic's x)
    int32_t base;       // input sample under tap 0 of phase `sub` (absolute; < 0 / >= n_in read as silence)
    int32_t sub;        // phase 0 .. 255; the cubic also reads sub - 1, sub + 1, sub + 2 (carries move `base` by one)
};

struct ResamplePlan {
    AfResamplerSpec spec{};
    size_t n_in = 0;
    AfResamplerShape shape{};
    double cutoff = 0.0;                 // after the down-sampling scale (make_interpolator)
    std::vector<double> table;           // [256][sinc_len]
    std::vector<ResampleFrame> frames;   // shape.frames entries
    int max_span = 0;                    // largest input span (samples) one CTA of kResampleFramesPerBlock frames stages
};

// Validation in the reference's order with its messages (resampling.rs:187-221); the configurations with a pinned
// `calculate_cutoff` value only (else AFSIM_UNSUPPORTED).  `with_table = false` skips table and frame list (shape only).
int plan_resampler(const AfResamplerSpec& spec, size_t n_in, bool with_frames, bool with_table, ResamplePlan* plan, std::string* msg);

// d_in: [n_streams][in_stride] f64, d_out: [n_streams][out_stride] f64; frames / table on the device.
cudaError_t launch_resample(const double* d_in, size_t in_stride, size_t n_in, double* d_out, size_t out_stride, size_t n_frames,
                            int n_streams, const ResampleFrame* d_frames, const double* d_table, int sinc_len, int max_span,
                            cudaStream_t stream);

}  // namespace afsim
