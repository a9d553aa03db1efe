"""Pins the resampler oracle (oracle/resampler_oracle.py: restated rubato 0.14 `SincFixedIn<f64>`, cubic, 256 phases)
on the study the REAL reference published: runs the reference's OWN tool code
(`python/tools/evaluate_resampler_quality.py`, `_evaluate_configuration`: pass-band tones against an offline
scipy reference, stop-band tones + swept noise, image tones, impulse location, the 44.1 -> 48 -> 44.1 kHz pink-noise
round trip, 60 s sample counts and delays) with the oracle standing in for `mic_eq_core.simulate_product_resampler`
(`mic_eq/__init__.py:38-73` picks up a top-level `mic_eq_core`), for all three configurations of
`evaluation/resampler-quality-report.json` -- product (Blackman 128), legacy (Blackman-Harris^2 128) and
high-rejection (Blackman-Harris^2 256).  The report records the hashes of `resampling.rs` and the tool: equal to this
tree's files (checked below).  Wall-clock timing keys are dropped.

Run in the build container (imports the reference tree, copies nothing).  Writes tests/golden/resampler_report.json:
the published values beside the oracle's; tests/test_oracle_resampler_report.py checks them."""
import hashlib
import json
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
from oracle import resampler_oracle  # noqa: E402

TIMING = ("block_time_ns", "deadline_ns", "p99_deadline_fraction", "max_deadline_fraction", "blocks")


def install_shim(simulate=None):
    shim = types.ModuleType("mic_eq_core")
    shim.AudioProcessor = type("AudioProcessor", (), {})
    shim.DeviceInfo = type("DeviceInfo", (), {})
    shim.list_input_devices = lambda: []
    shim.list_output_devices = lambda: []

    def door(samples, input_rate, output_rate, chunk_size=1024, sinc_len=None, window=None):
        out, delay, expected, _ = (simulate or resampler_oracle.simulate_product_resampler)(
            samples, input_rate, output_rate, chunk_size, sinc_len, window)
        return out, delay, expected, [1]

    shim.simulate_product_resampler = door
    shim.product_resampler_configuration = resampler_oracle.product_resampler_configuration
    sys.modules["mic_eq_core"] = shim
    sys.path.insert(0, str(REF / "python"))
    sys.path.insert(0, str(REF / "python" / "tools"))


def leaves(node, prefix=""):
    if isinstance(node, dict):
        for k, v in node.items():
            if k in TIMING:
                continue
            yield from leaves(v, f"{prefix}/{k}")
    elif isinstance(node, list):
        for i, v in enumerate(node):
            yield from leaves(v, f"{prefix}[{i}]")
    else:
        yield prefix, node


def main():
    published = json.loads((REF / "evaluation" / "resampler-quality-report.json").read_text())
    for rel, want in published["source_sha256"].items():
        got = hashlib.sha256((REF / rel).read_bytes()).hexdigest()
        assert got == want, f"{rel}: the report was produced from another source version"
    install_shim()
    import evaluate_resampler_quality as tool
    duration = published["configuration"]["long_stream_duration_seconds"]
    configs = {
        "product": tool.ResamplerConfiguration("product", 128, "blackman", native_default=True),
        "legacy-blackman-harris-squared-128": tool.ResamplerConfiguration("legacy-blackman-harris-squared-128", 128, "blackman_harris_squared"),
        "high-rejection-blackman-harris-squared-256": tool.ResamplerConfiguration("high-rejection-blackman-harris-squared-256", 256, "blackman_harris_squared"),
    }
    pub = {"product": published["product"]}
    pub.update({a["configuration"]["identifier"]: a for a in published["alternatives"]})
    golden = {"source_sha256": published["source_sha256"], "cutoffs": {f"{k[0]}:{k[1]}": v for k, v in resampler_oracle.KNOWN_CUTOFFS.items()},
              "configurations": {}}
    worst = 0.0
    for name, cfg in configs.items():
        mine = tool._evaluate_configuration(cfg, duration)
        rows = {}
        want = dict(leaves({k: pub[name][k] for k in ("status", "checks", "measurements")}))
        got = dict(leaves({k: mine[k] for k in ("status", "checks", "measurements")}))
        # the published report keeps the summary of each pass-band / alias / image case; compare what it holds
        for key, w in want.items():
            g = got[key]
            rows[key] = {"published": w, "oracle": g}
            if isinstance(w, float) and not isinstance(w, bool):
                worst = max(worst, abs(g - w))
                print(f"{name}{key}: {w!r} {g!r} {g - w:+.3e}")
            else:
                assert g == w, (name, key, w, g)
        golden["configurations"][name] = rows
    golden["largest_absolute_difference"] = worst
    out = ROOT / "tests" / "golden" / "resampler_report.json"
    out.write_text(json.dumps(golden, indent=1, sort_keys=True) + "\n")
    print("wrote", out, "largest |difference|", worst)


if __name__ == "__main__":
    main()
