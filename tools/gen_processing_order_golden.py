"""Pins the CPU oracle on numbers the REAL reference produced: the EQ / de-esser order study of
`python/tools/evaluate_processing_order.py:369-507`, whose published report
(`evaluation/processing-order-report.json`, section `eq_deesser`, native_simulation_backend = mic_eq_core, source
hashes equal to this reference tree) holds six medians over the reference's own 96-case generated de-esser corpus
(`python/mic_eq/analysis/deesser_corpus.py`, numpy only, CC0): de-esser peak reduction of the clips that need /
do not need de-essing with the de-esser before and after the EQ, and the 4 - 10 kHz energy change of the "bright"
clips computed from the returned audio.  Every clip is rendered twice through `simulate_auto_eq_chain` (legacy
bands with the 72-sample fade-in, de-esser auto mode 0.75, compressor and limiter off, both orders, 44.1 and 48 kHz).

Run in the build container (the reference tree is imported for the corpus generator and the tool's own
`_band_energy`; nothing of it is copied): regenerates the corpus, renders it with the ORACLE as the native door,
recomputes the six medians exactly as the tool does and writes tests/golden/processing_order.json with the
published values, the oracle's values and the per-clip rows.  tests/test_oracle_reference_report.py checks them.
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF / "python"))

from audio_forge_b200 import abi, mic_eq_core  # noqa: E402
from oracle import pyoracle  # noqa: E402

BANDS = [(80.0, 0.0, 1.0), (140.0, 0.0, 1.0), (250.0, 0.0, 1.0), (450.0, 0.0, 1.0), (800.0, 0.0, 1.0), (1_500.0, 0.0, 1.0),
         (2_800.0, 1.5, 1.0), (5_000.0, 4.0, 1.1), (8_000.0, 3.0, 1.0), (12_000.0, 0.0, 1.0)]  # evaluate_processing_order.py:373-384
COMMON = {"deesser_enabled": True, "deesser_auto_enabled": True, "deesser_auto_amount": 0.75, "compressor_enabled": False,
          "limiter_enabled": False, "return_output_audio": True}  # :385-392


def oracle_door(audio, sample_rate, bands, settings):
    """simulate_auto_eq_chain with the oracle behind it (same settings parsing as the product's Python door)."""
    st, _, _ = mic_eq_core.settings_from_mapping(settings)
    m, out, _ = pyoracle.chain_render(np.asarray(audio, dtype=np.float32), float(sample_rate), abi.legacy_bands(bands), st,
                                      return_audio=True)
    result = abi.metrics_to_dict(m)
    result["output_audio"] = out
    return result


def study(door):
    import importlib

    from mic_eq.analysis.deesser_corpus import CORPUS_CASES, generate_deesser_case

    sys.path.insert(0, str(REF / "python" / "tools"))
    band_energy = importlib.import_module("evaluate_processing_order")._band_energy  # the tool's own measure (:362-366)

    rows = []
    for spec in CORPUS_CASES:
        generated = generate_deesser_case(spec)
        high = min(10_000.0, spec.sample_rate * 0.45)
        baseline = door(generated.speech_audio, spec.sample_rate, BANDS, {**COMMON, "eq_before_deesser": False})
        candidate = door(generated.speech_audio, spec.sample_rate, BANDS, {**COMMON, "eq_before_deesser": True})
        input_hf = band_energy(generated.speech_audio.astype(np.float64), spec.sample_rate, 4_000.0, high)
        row = {"id": spec.name, "needs_deesser": bool(spec.needs_deesser), "condition": spec.condition,
               "sample_rate": int(spec.sample_rate),
               "input_sha256": hashlib.sha256(np.ascontiguousarray(generated.speech_audio, dtype=np.float32).tobytes()).hexdigest()[:16],
               "baseline_peak_reduction_db": float(baseline["deesser_gain_reduction_db"]),
               "candidate_peak_reduction_db": float(candidate["deesser_gain_reduction_db"])}
        for name, sim in (("baseline", baseline), ("candidate", candidate)):
            audio = np.asarray(sim["output_audio"], dtype=np.float64)
            row[f"{name}_hf_change_db"] = 10.0 * np.log10(max(band_energy(audio, spec.sample_rate, 4_000.0, high), 1e-18)
                                                          / max(input_hf, 1e-18))
        rows.append(row)
    positive = [r for r in rows if r["needs_deesser"]]
    negative = [r for r in rows if not r["needs_deesser"]]
    bright = [r for r in negative if r["condition"] == "bright"]
    med = lambda group, key: float(np.median([r[key] for r in group]))  # noqa: E731
    metrics = {"positive_baseline_peak_reduction_db": med(positive, "baseline_peak_reduction_db"),
               "positive_candidate_peak_reduction_db": med(positive, "candidate_peak_reduction_db"),
               "negative_baseline_peak_reduction_db": med(negative, "baseline_peak_reduction_db"),
               "negative_candidate_peak_reduction_db": med(negative, "candidate_peak_reduction_db"),
               "bright_baseline_hf_change_db": med(bright, "baseline_hf_change_db"),
               "bright_candidate_hf_change_db": med(bright, "candidate_hf_change_db")}
    return metrics, rows


def main():
    report = json.loads((REF / "evaluation" / "processing-order-report.json").read_text())
    published = report["eq_deesser"]["metrics"]
    hashes_ok = all(hashlib.sha256((REF / rel).read_bytes()).hexdigest() == digest
                    for rel, digest in report["provenance"]["source_hashes"].items() if (REF / rel).exists())
    metrics, rows = study(oracle_door)
    out = {"source": "evaluation/processing-order-report.json (eq_deesser), produced by the reference's native core",
           "report_source_hashes_match_this_tree": bool(hashes_ok), "published": published, "oracle": metrics, "cases": rows}
    (ROOT / "tests" / "golden" / "processing_order.json").write_text(json.dumps(out, indent=1) + "\n")
    for key in published:
        print(f"{key:45s} published {published[key]!r:24} oracle {metrics[key]!r:24} diff {metrics[key] - published[key]:+.3e}")
    print("source hashes match:", hashes_ok, "cases:", len(rows))


if __name__ == "__main__":
    main()
