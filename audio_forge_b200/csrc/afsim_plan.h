// afsim_plan.h -- host planner: candidate settings -> the lane tables the kernels read.
//
// This is the host side of the drop-in: it plays the role of the reference's constructor +
// setter sequence (rust-core/src/audio/processor/python_api.rs:400-487) and reduces it to the
// constants each recurrence needs, using the host libm exactly as the reference does (so the
// coefficients are bit-identical to a CPU render), then packs 32 streams per warp.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/afsim.h"
#include "afsim_layout.h"

namespace afsim {

struct BiquadCoeffs {
    double b0, b1, b2, a1, a2;
};
enum class BqKind { LowShelf, HighShelf, Peaking, Notch, HighPass, LowPass };

BiquadCoeffs design_biquad(BqKind kind, double frequency_hz, double gain_db, double q, double sample_rate);
double time_constant_to_coeff(double time_ms, double sample_rate);

// Everything one stream needs, before packing.
struct StreamPlan {
    double params[P_COUNT];
    float fparams[FP_COUNT];
    uint32_t lane_flags;
    uint32_t n_sections;
    uint32_t group_flags;     // structural flags every lane of a warp must share
    uint32_t lookahead;
    float effective_ceiling_db;
    int band_sections[AFSIM_NUM_BANDS];  // sections per band (for the response renderer)
};

struct RateConstants {
    uint32_t block_samples;
    uint32_t fade_samples;
    double eq_default[10][5];
};

// Returns "" on success, else the reference's validation message (status in *status).
std::string plan_stream(const AfBand bands[AFSIM_NUM_BANDS], const AfChainSettings& settings, double sample_rate,
                        StreamPlan* out, int* status);
std::string validate_typed_bands(const AfBand bands[AFSIM_NUM_BANDS], double sample_rate);
std::string validate_legacy_response_bands(const AfBand bands[AFSIM_NUM_BANDS], double sample_rate);
RateConstants rate_constants(double sample_rate);

// EQ sections only (for afsim_eq_response): coefficient table [40][5] in band*4+section order.
void plan_eq_sections(const AfBand bands[AFSIM_NUM_BANDS], bool typed, double sample_rate, double coeffs[kMaxSections][5],
                      int band_sections[AFSIM_NUM_BANDS]);

// Packing: streams -> warps.
struct PackedStream {
    uint32_t pair;       // caller's pair index
    uint32_t passage;
    uint32_t candidate;
};
struct PackedGroup {
    GroupHeader header;
    uint32_t pair_of_lane[kLanes];  // 0xffffffff for padding lanes
    uint32_t plan_of_lane[kLanes];  // candidate index whose plan the lane uses
};

}  // namespace afsim
