"""Pins the CPU oracle on two more studies the REAL reference published, by running the reference's OWN tool code
with the oracle standing in for the native core (`mic_eq/__init__.py:38-46` picks up a top-level `mic_eq_core`):

* `python/tools/evaluate_limiter_lookahead.py` -> `evaluation/limiter-lookahead-report.json`, the "controlled"
  aggregates (three generated fixtures: sine bursts, impulses, clipped voice; compressor + lookahead limiter at
  0.5 / 1 / 2 ... ms + true-peak limiter + detector): true-peak limiter gain reduction, limited events, pre / output
  true-peak overshoot, gain-envelope variation and transient shape error of the returned audio.
* `python/tools/evaluate_dynamics_aliasing.py` -> `evaluation/dynamics-aliasing-report.json`: four AM carriers through
  the compressor (0.5 ms attack, 8:1) at 48 kHz and 192 kHz: peak gain reduction at both rates, alignment lag,
  waveform / folded error of the 48 kHz render against the decimated 192 kHz render.

* `python/tools/evaluate_eq_filter_types.py` -> `evaluation/eq-filter-types-report.json`, analytic measurements
  and headroom prediction: `eq_magnitude_response(_v2)` on the default bands, every Butterworth slope at its cutoff,
  a notch, 250 random typed 10-band settings, and `simulate_eq_v2` on a 2 s sine through a 12 dB bell.

Run in the build container (imports the reference tree, copies nothing).  Writes tests/golden/reference_reports.json:
the published values beside the oracle's; tests/test_oracle_reference_report.py checks them.
"""
import importlib
import json
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

from audio_forge_b200 import abi, mic_eq_core as product_door  # noqa: E402
from oracle import pyoracle  # noqa: E402


def oracle_simulate_auto_eq_chain(audio, sample_rate, bands, settings=None):
    st, typed, return_audio = product_door.settings_from_mapping(settings)
    band_arr = abi.typed_bands(typed) if typed is not None else abi.legacy_bands(bands)
    m, out, _ = pyoracle.chain_render(np.asarray(audio, dtype=np.float32), float(sample_rate), band_arr, st,
                                      return_audio=return_audio)
    result = abi.metrics_to_dict(m)
    if return_audio:
        result["output_audio"] = out
    return result


def oracle_eq_magnitude_response(frequencies_hz, bands, sample_rate):
    return pyoracle.eq_response(list(frequencies_hz), abi.legacy_bands(bands), float(sample_rate), typed=False).tolist()


def oracle_eq_magnitude_response_v2(frequencies_hz, bands, sample_rate):
    return pyoracle.eq_response(list(frequencies_hz), abi.typed_bands(bands), float(sample_rate), typed=True).tolist()


def oracle_simulate_eq_v2(audio, sample_rate, bands, return_output_audio=False):
    st, out = pyoracle.eq_render(np.asarray(audio, dtype=np.float32), float(sample_rate), abi.typed_bands(bands),
                                 return_audio=return_output_audio)
    result = {name: getattr(st, name) for name, _ in st._fields_}
    if return_output_audio:
        result["output_audio"] = out
    return result


def install_shim():
    shim = types.ModuleType("mic_eq_core")
    shim.AudioProcessor = type("AudioProcessor", (), {})
    shim.DeviceInfo = type("DeviceInfo", (), {})
    shim.list_input_devices = lambda: []
    shim.list_output_devices = lambda: []
    shim.simulate_auto_eq_chain = oracle_simulate_auto_eq_chain
    shim.eq_magnitude_response = oracle_eq_magnitude_response
    shim.eq_magnitude_response_v2 = oracle_eq_magnitude_response_v2
    shim.simulate_eq_v2 = oracle_simulate_eq_v2
    sys.modules["mic_eq_core"] = shim
    sys.path.insert(0, str(REF / "python"))
    sys.path.insert(0, str(REF / "python" / "tools"))


RUNTIME_KEYS = ("runtime", "generated", "environment")


def limiter_study():
    tool = importlib.import_module("evaluate_limiter_lookahead")
    tool.RUNTIME_REPETITIONS = 1  # the repetitions only feed the runtime statistics
    report = json.loads((REF / "evaluation" / "limiter-lookahead-report.json").read_text())
    fixtures = tool._cases()
    out = {}
    for key, groups in report["aggregates"].items():
        lookahead = float(key)
        rows = [tool._case(name, audio, lookahead) for name, audio in sorted(fixtures.items())]
        ours = tool._aggregate(rows)
        published = groups["controlled"]
        keep = [k for k in published if "runtime" not in k]
        out[key] = {"published": {k: published[k] for k in keep}, "oracle": {k: ours[k] for k in keep}}
    return out


def dynamics_study():
    tool = importlib.import_module("evaluate_dynamics_aliasing")
    report = json.loads((REF / "evaluation" / "dynamics-aliasing-report.json").read_text())
    out = []
    for published in report["cases"]:
        ours = tool._case(published["id"], published["carrier_hz"], published["modulation_hz"])
        keep = [k for k in published if "runtime" not in k]
        out.append({"published": {k: published[k] for k in keep}, "oracle": {k: ours[k] for k in keep}})
    return out


def eq_filter_types_study():
    """evaluate_eq_filter_types.py: the analytic measurements (response renderer: default bands legacy vs typed, the
    -3.01 dB cutoff of every Butterworth slope, the notch, 250 random 10-band settings from rng 0xE041) and the
    headroom prediction (a 12 dB bell at 1 kHz: response renderer against simulate_eq_v2 on a 2 s sine)."""
    tool = importlib.import_module("evaluate_eq_filter_types")
    report = json.loads((REF / "evaluation" / "eq-filter-types-report.json").read_text())["measurements"]
    published = {"analytic": report["analytic"], "headroom_prediction": report["headroom_prediction"]}
    cases = int(report["analytic"]["random_boundary_stress"]["cases"])
    ours = {"analytic": tool._analytic_measurements(cases), "headroom_prediction": tool._headroom_prediction_measurement()}
    return {"published": published, "oracle": ours}


def source_hashes_match():
    """The reports that record hashes of the sources they were produced from: are those this tree's files?"""
    import hashlib
    out = {}
    for name in ("limiter-lookahead-report.json", "eq-filter-types-report.json"):
        hashes = json.loads((REF / "evaluation" / name).read_text())["source_sha256"]
        out[name] = all((REF / rel).exists() and hashlib.sha256((REF / rel).read_bytes()).hexdigest() == digest
                        for rel, digest in hashes.items())
        out[name + ":files"] = sorted(hashes)
    return out


def _flatten(obj, prefix=""):
    if isinstance(obj, dict):
        for k, v in obj.items():
            yield from _flatten(v, f"{prefix}{k}.")
    elif isinstance(obj, list):
        for i, v in enumerate(obj):
            yield from _flatten(v, f"{prefix}{i}.")
    else:
        yield prefix[:-1], obj


def main():
    install_shim()
    result = {"limiter_lookahead_controlled": limiter_study(), "dynamics_aliasing": dynamics_study(),
              "eq_filter_types": eq_filter_types_study(), "report_source_hashes": source_hashes_match()}
    pub, ours = dict(_flatten(result["eq_filter_types"]["published"])), dict(_flatten(result["eq_filter_types"]["oracle"]))
    for k, v in pub.items():
        if isinstance(v, str):
            continue
        print(f"eq types {k:58s} published {v!r:24} oracle {ours.get(k)!r:24} |diff| {abs(float(ours[k]) - float(v)):.3e}")
    (ROOT / "tests" / "golden" / "reference_reports.json").write_text(json.dumps(result, indent=1) + "\n")
    worst = 0.0
    for key, entry in result["limiter_lookahead_controlled"].items():
        for k, v in entry["published"].items():
            o = entry["oracle"][k]
            d = abs(float(o) - float(v)) if not isinstance(v, bool) else float(o != v)
            worst = max(worst, d)
            print(f"limiter {key:4s} {k:45s} published {v!r:24} oracle {o!r:24} |diff| {d:.3e}")
    for entry in result["dynamics_aliasing"]:
        for k, v in entry["published"].items():
            if isinstance(v, str):
                continue
            o = entry["oracle"][k]
            d = abs(float(o) - float(v))
            worst = max(worst, d)
            print(f"dynamics {entry['published']['id']:12s} {k:36s} published {v!r:24} oracle {o!r:24} |diff| {d:.3e}")
    print("worst |diff|", worst)


if __name__ == "__main__":
    main()
