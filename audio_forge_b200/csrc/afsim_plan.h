// afsim_plan.h -- host planner: candidate settings -> the constants the kernels read.
//
// This is the host side of the drop-in: it plays the role of the reference's constructor +
// setter sequence (rust-core/src/audio/processor/python_api.rs:400-487) and reduces it to the
// constants each recurrence needs, using the host libm exactly as the reference does (so the
// coefficients are bit-identical to a CPU render).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/afsim.h"
#include "afsim_cleanup.h"
#include "afsim_params.h"

namespace afsim {

struct BiquadCoeffs {
    double b0, b1, b2, a1, a2;
};
enum class BqKind { LowShelf, HighShelf, Peaking, Notch, HighPass, LowPass };

BiquadCoeffs design_biquad(BqKind kind, double frequency_hz, double gain_db, double q, double sample_rate);
double time_constant_to_coeff(double time_ms, double sample_rate);

struct CandidatePlan {
    CandidateParams params;
    uint32_t structure;   // StructureFlag
    uint32_t lookahead;   // limiter lookahead in samples (0 when the limiter is off)
    uint32_t input_stage; // AfInputStage
    // the de-esser's configured band edges reach Nyquist (e.g. the 11 kHz default at 16 kHz): its detector / dynamic
    // EQ biquads are designed with w0 >= pi and are unstable in the reference itself.  The library rejects such a
    // render (AFSIM_UNSUPPORTED) instead of returning numbers that amplify 1-ulp libm differences without bound;
    // tests/hostsim ignores the flag to pin the non-finite semantics against the oracle.
    uint32_t deesser_unstable;
};

struct RateConstants {
    int block_samples;            // python_api.rs:512-513
    int fade_samples;             // dsp/biquad.rs:12-19
    double eq_default[10][5];     // constructor coefficients of the 10 default bands (dsp/eq.rs:125-140)
    CleanupConst cleanup;         // adaptive input cleanup constants (audio/processor/routing.rs:55-330)
};

// Fills `out`; returns AFSIM_OK or a status with the reference's message in *error.
int plan_candidate(const AfBand bands[AFSIM_NUM_BANDS], const AfChainSettings& settings, double sample_rate,
                   CandidatePlan* out, std::string* error);
// simulate_eq_v2 (lib.rs:214-288): typed bands, EQ only, true-peak detectors on both sides.
int plan_eq_only(const AfBand bands[AFSIM_NUM_BANDS], double sample_rate, CandidatePlan* out, std::string* error);
int validate_typed_bands(const AfBand bands[AFSIM_NUM_BANDS], double sample_rate, std::string* error);
int validate_legacy_response_bands(const AfBand bands[AFSIM_NUM_BANDS], double sample_rate, std::string* error);
int validate_response_frequencies(const double* freqs, size_t n, double sample_rate, std::string* error);
RateConstants rate_constants(double sample_rate);
// Loudness meter of the compressor's auto makeup for this rate, block length and render length.
MakeupConst makeup_constants(double sample_rate, int block_samples, int n_samples);
// simulate_auto_makeup_control (python_api.rs:118-276).
int plan_makeup_control(const AfAutoMakeupSettings& settings, double sample_rate, double noise_floor_db,
                        double noise_reliability, bool has_vad, CandidatePlan* out, std::string* error);

// EQ sections only (afsim_eq_response): coefficient table [40][5] in band*4+section order and the
// number of sections per band.
void plan_eq_sections(const AfBand bands[AFSIM_NUM_BANDS], bool typed, double sample_rate, double coeffs[kMaxSections][5],
                      int band_sections[AFSIM_NUM_BANDS]);

// Large compressor grids (afsim_api.cu, build_sweep): the streams of one batch (`passage[i]`, `eq_class[i]` per
// stream, any order) are cut into pieces of at most `max_streams` when they cover few distinct (passage, EQ class)
// pairs (4 x pairs <= max_streams, so that every piece still shares its prefix).  The streams of a passage stay
// together (stable order otherwise), pieces are balanced and a multiple of 32 streams long (the last one may be
// shorter).  Returns the pieces as lists of stream positions; one piece = not cut.
std::vector<std::vector<uint32_t>> cut_stream_group(const std::vector<uint32_t>& passage, const std::vector<uint32_t>& eq_class,
                                                    int max_streams);

void chain_settings_default(AfChainSettings* out);
void default_bands(AfBand out[AFSIM_NUM_BANDS]);

}  // namespace afsim
