"""Pins the CPU oracle on two more studies the REAL reference published, by running the reference's OWN tool code
with the oracle standing in for the native core (`mic_eq/__init__.py:38-46` picks up a top-level `mic_eq_core`):

* `python/tools/evaluate_limiter_lookahead.py` -> `evaluation/limiter-lookahead-report.json`, the "controlled"
  aggregates (three generated fixtures: sine bursts, impulses, clipped voice; compressor + lookahead limiter at
  0.5 / 1 / 2 ... ms + true-peak limiter + detector): true-peak limiter gain reduction, limited events, pre / output
  true-peak overshoot, gain-envelope variation and transient shape error of the returned audio.
* `python/tools/evaluate_dynamics_aliasing.py` -> `evaluation/dynamics-aliasing-report.json`: four AM carriers through
  the compressor (0.5 ms attack, 8:1) at 48 kHz and 192 kHz: peak gain reduction at both rates, alignment lag,
  waveform / folded error of the 48 kHz render against the decimated 192 kHz render.

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


def install_shim():
    shim = types.ModuleType("mic_eq_core")
    shim.AudioProcessor = type("AudioProcessor", (), {})
    shim.DeviceInfo = type("DeviceInfo", (), {})
    shim.list_input_devices = lambda: []
    shim.list_output_devices = lambda: []
    shim.simulate_auto_eq_chain = oracle_simulate_auto_eq_chain
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


def main():
    install_shim()
    result = {"limiter_lookahead_controlled": limiter_study(), "dynamics_aliasing": dynamics_study()}
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
