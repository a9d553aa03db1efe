"""Golden decisions for the compressor calibration search, produced by the REFERENCE's own search code.

Runs `mic_eq.analysis.voice_setup._calibrate_compressor_threshold` (imported from /root/reference, pure
Python) with its native door `simulate_candidate_chain` replaced by the CPU oracle, on synthetic captures,
and writes the selected settings + objective values to tests/golden/compressor_search.json.  Run in the
build container only (the reference tree is not on the GPU box); the fixture is committed.

    PYTHONPATH=/root/reference/python python tools/gen_compressor_search_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/python")

from audio_forge_b200 import abi, headroom, mic_eq_core  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests.signals import speech_like  # noqa: E402

from mic_eq.analysis import voice_setup  # noqa: E402  (the reference)

CASES = [
    dict(name="default_incumbent", seed=11, seconds=3.0, level=0.5,
         compressor=dict(enabled=True, threshold_db=-20.0, ratio=4.0, attack_ms=10.0, release_ms=200.0, makeup_gain_db=0.0,
                         adaptive_release=True, base_release_ms=50.0, auto_makeup_enabled=False, target_lufs=-18.0,
                         sidechain_highpass_enabled=True, measured_short_term_lufs=-23.0),
         targets=dict(target_p95_db=6.0, target_median_db=3.0, peak_cap_db=12.0)),
    dict(name="hot_capture_tight_cap", seed=12, seconds=3.0, level=0.9,
         compressor=dict(enabled=True, threshold_db=-30.0, ratio=2.5, attack_ms=5.0, release_ms=120.0, makeup_gain_db=0.0,
                         adaptive_release=False, base_release_ms=50.0, auto_makeup_enabled=True, target_lufs=-16.0,
                         sidechain_highpass_enabled=True, measured_short_term_lufs=-17.0),
         targets=dict(target_p95_db=4.5, target_median_db=2.0, peak_cap_db=8.0)),
    dict(name="quiet_capture", seed=13, seconds=3.0, level=0.12,
         compressor=dict(enabled=True, threshold_db=-40.0, ratio=3.0, attack_ms=15.0, release_ms=250.0, makeup_gain_db=0.0,
                         adaptive_release=True, base_release_ms=50.0, auto_makeup_enabled=False, target_lufs=-18.0,
                         sidechain_highpass_enabled=False, measured_short_term_lufs=-36.0),
         targets=dict(target_p95_db=5.0, target_median_db=2.5, peak_cap_db=10.0)),
]
EQ = {"band_freqs": list(abi.DEFAULT_FREQUENCIES), "band_gains": [0.0, 1.5, 0.0, -2.0, 0.0, 2.0, 3.0, 0.0, 1.0, 0.0],
      "band_qs": [1.41] * 10}
DEESSER = {"enabled": False}


def oracle_simulate_candidate_chain(audio, sample_rate, eq_settings, chain_settings=None):
    """headroom.simulate_candidate_chain with the oracle standing in for the native core."""
    flat = headroom.flatten_chain_settings(chain_settings)
    flat.pop("return_output_audio")
    st, _, _ = mic_eq_core.settings_from_mapping(flat)
    m, _, _ = pyoracle.chain_render(np.ascontiguousarray(audio, dtype=np.float32), float(sample_rate),
                                    abi.legacy_bands(headroom.bands_from_settings(eq_settings)), st)
    out = abi.metrics_to_dict(m)
    out["simulation_backend"] = "rust"
    out["safety_authority"] = "authoritative"
    return out


def main():
    voice_setup.simulate_candidate_chain = oracle_simulate_candidate_chain
    golden = []
    for case in CASES:
        audio = speech_like(int(case["seconds"] * 48000), seed=case["seed"], level=case["level"])
        calibrated, diag = voice_setup._calibrate_compressor_threshold(
            speech_audio=audio, sample_rate=48000, eq_settings=EQ, deesser_settings=DEESSER,
            compressor_settings=case["compressor"], **case["targets"])
        golden.append({
            "case": case,
            "selected": {k: calibrated[k] for k in ("threshold_db", "ratio", "attack_ms", "release_ms")},
            "iterations": diag["iterations"],
            "expanded_search_selected": diag["expanded_search_selected"],
            "total_objective": diag["total_objective"],
            "threshold_only_objective": diag["threshold_only_objective"],
            "expanded_candidate_objective": diag["expanded_candidate_objective"],
            "incumbent_objective": diag["incumbent_objective"],
        })
        print(case["name"], golden[-1]["selected"], diag["iterations"], diag["expanded_search_selected"], diag["total_objective"])
    out = {"eq_settings": EQ, "deesser_settings": DEESSER, "cases": golden,
           "generator": "tools/gen_compressor_search_golden.py (reference voice_setup._calibrate_compressor_threshold + CPU oracle)"}
    path = ROOT / "tests" / "golden" / "compressor_search.json"
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(json.dumps(out, indent=1))
    print("wrote", path)


if __name__ == "__main__":
    main()
