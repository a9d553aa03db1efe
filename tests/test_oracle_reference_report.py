"""The oracle against numbers the REAL reference published (evaluation/*.json of the reference tree, produced by its
native core `mic_eq_core`): the EQ / de-esser order study over the reference's 96-case generated corpus, the
controlled fixtures of the limiter-lookahead study and the 48 / 192 kHz dynamics-aliasing study.

tests/golden/processing_order.json and tests/golden/reference_reports.json hold the published values beside what
the oracle produced when the reference's own tool code ran with the oracle as its native core
(tools/gen_processing_order_golden.py, tools/gen_reference_report_golden.py).  The committed pairs are checked
always; when the reference tree is present (the build container) part of each study is recomputed live.
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).parent / "golden"
REF = Path("/root/reference")
F32_ULP_DB = 1.2e-7  # one f32 ulp at 0.5 - 1 dB: the reference's f32 log10 is the Windows CRT's, the oracle's glibc's


def _close(key, published, ours):
    if isinstance(published, (bool, str)) or isinstance(published, int):
        return published == ours
    if key in ("max_true_peak_limiter_gain_reduction_db", "worst_pre_true_peak_overshoot_db"):
        return abs(published - ours) <= F32_ULP_DB  # f32 dB conversions (python_api.rs:54-56)
    if key in ("relative_waveform_error_db", "folded_out_of_expected_error_db"):
        return abs(published - ours) <= 1e-9  # scipy resample_poly / FFT rounding of the tool itself
    return published == ours


def test_processing_order_study_matches_the_published_medians_bit_for_bit():
    g = json.loads((GOLDEN / "processing_order.json").read_text())
    assert g["report_source_hashes_match_this_tree"] is True
    assert len(g["cases"]) == 96 and {c["sample_rate"] for c in g["cases"]} == {44100, 48000}
    for key, published in g["published"].items():
        assert g["oracle"][key] == published, key  # six medians, incl. two computed from the returned audio
    assert g["published"]["negative_candidate_peak_reduction_db"] == 0.8453983068466187


def test_limiter_and_dynamics_studies_match_the_published_values():
    g = json.loads((GOLDEN / "reference_reports.json").read_text())
    # the reports were produced from exactly the sources of this reference tree (limiter.rs, true_peak.rs, biquad.rs,
    # eq.rs, lib.rs, python_api.rs and the tools themselves)
    assert g["report_source_hashes"]["limiter-lookahead-report.json"] is True
    assert g["report_source_hashes"]["eq-filter-types-report.json"] is True
    assert "rust-core/src/dsp/true_peak.rs" in g["report_source_hashes"]["limiter-lookahead-report.json:files"]
    assert sorted(g["limiter_lookahead_controlled"]) == ["0.5", "1.0", "2.0"]
    exact = 0
    for key, entry in g["limiter_lookahead_controlled"].items():
        for k, published in entry["published"].items():
            assert _close(k, published, entry["oracle"][k]), (key, k)
            exact += published == entry["oracle"][k]
    assert exact >= 27  # everything but the two f32 dB conversions per lookahead is identical
    assert len(g["dynamics_aliasing"]) == 4
    for entry in g["dynamics_aliasing"]:
        for k, published in entry["published"].items():
            assert _close(k, published, entry["oracle"][k]), (entry["published"]["id"], k)
        # the compressor's peak gain reduction at 48 kHz and at 192 kHz: identical f32 values
        assert entry["oracle"]["base_peak_gain_reduction_db"] == entry["published"]["base_peak_gain_reduction_db"]
        assert entry["oracle"]["reference_peak_gain_reduction_db"] == entry["published"]["reference_peak_gain_reduction_db"]


def test_eq_filter_types_study_matches_the_published_values():
    """Response renderer and simulate_eq_v2 (evaluation/eq-filter-types-report.json, analytic + headroom prediction):
    notch probes, the maximum over 250 random 10-band typed settings (rng 0xE041) and the measured gain of a 12 dB
    bell through simulate_eq_v2 identical; the Butterworth cutoffs within 2e-15 dB (libm ulps of another platform)."""
    g = json.loads((GOLDEN / "reference_reports.json").read_text())["eq_filter_types"]
    pub, ours = g["published"], g["oracle"]
    assert ours["analytic"]["notch"]["response_db"] == pub["analytic"]["notch"]["response_db"]
    assert ours["analytic"]["random_boundary_stress"] == pub["analytic"]["random_boundary_stress"]
    assert pub["analytic"]["random_boundary_stress"]["max_absolute_response_db"] == 1206.650162779802
    assert ours["analytic"]["default_response_max_absolute_delta_db"] == pub["analytic"]["default_response_max_absolute_delta_db"] == 0.0
    for a, b in zip(pub["analytic"]["cutoff"], ours["analytic"]["cutoff"]):
        assert (a["filter_type"], a["slope_db_per_octave"]) == (b["filter_type"], b["slope_db_per_octave"])
        assert abs(a["measured_db"] - b["measured_db"]) <= 4e-15
    for key in ("predicted_gain_db", "measured_gain_db", "absolute_error_db", "reported_max_response_db"):
        assert ours["headroom_prediction"][key] == pub["headroom_prediction"][key], key


@pytest.mark.skipif(not (REF / "python" / "tools" / "evaluate_dynamics_aliasing.py").exists(),
                    reason="reference tree not present (only in the build container)")
def test_live_recomputation_with_the_reference_tool_code():
    """The reference's own tool functions with the oracle as native core, now: one dynamics case (both rates), one
    limiter lookahead (three fixtures) and eight corpus clips of the order study."""
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from tools import gen_processing_order_golden as order
    from tools import gen_reference_report_golden as reports

    reports.install_shim()
    import importlib

    dyn = importlib.import_module("evaluate_dynamics_aliasing")
    g = json.loads((GOLDEN / "reference_reports.json").read_text())
    published = g["dynamics_aliasing"][1]["published"]
    ours = dyn._case(published["id"], published["carrier_hz"], published["modulation_hz"])
    for k, v in published.items():
        assert _close(k, v, ours[k]), k
    lim = importlib.import_module("evaluate_limiter_lookahead")
    lim.RUNTIME_REPETITIONS = 1
    rows = [lim._case(name, audio, 2.0) for name, audio in sorted(lim._cases().items())]
    agg = lim._aggregate(rows)
    for k, v in g["limiter_lookahead_controlled"]["2.0"]["published"].items():
        assert _close(k, v, agg[k]), k

    from mic_eq.analysis.deesser_corpus import CORPUS_CASES, generate_deesser_case
    committed = {c["id"]: c for c in json.loads((GOLDEN / "processing_order.json").read_text())["cases"]}
    for spec in list(CORPUS_CASES)[::12]:
        audio = generate_deesser_case(spec).speech_audio
        sim = order.oracle_door(audio, spec.sample_rate, order.BANDS, {**order.COMMON, "eq_before_deesser": True})
        assert float(sim["deesser_gain_reduction_db"]) == committed[spec.name]["candidate_peak_reduction_db"], spec.name
