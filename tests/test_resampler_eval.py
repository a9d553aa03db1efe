"""`audio_forge_b200.resampler_eval.evaluate_configuration` -- the reference's resampler-quality study in 11 batched native
calls instead of 28 -- with the CPU oracle behind the batch door, against the numbers the reference PUBLISHED with the real
crate (tests/golden/resampler_report.json holds them; evaluation/resampler-quality-report.json)."""
import json
from pathlib import Path

import numpy as np

from audio_forge_b200 import resampler_eval
from oracle import resampler_oracle as R

GOLDEN = json.loads((Path(__file__).parent / "golden" / "resampler_report.json").read_text())["configurations"]


def _oracle_batch(signals, input_rate, output_rate, chunk_size=1024, sinc_len=None, window=None):
    rows, delay, expected = [], 0, 0
    for row in np.asarray(signals, dtype=np.float64):
        out, delay, expected, _ = R.simulate_product_resampler(row, input_rate, output_rate, chunk_size, sinc_len, window)
        rows.append(out)
    return np.stack(rows), delay, expected


def _leaves(node, prefix=""):
    if isinstance(node, dict):
        for k, v in node.items():
            yield from _leaves(v, f"{prefix}/{k}")
    elif isinstance(node, list):
        for i, v in enumerate(node):
            yield from _leaves(v, f"{prefix}[{i}]")
    else:
        yield prefix, node


def test_batched_study_reproduces_the_published_product_configuration():
    got = resampler_eval.evaluate_configuration("product", 128, "blackman", native_default=True, duration_seconds=60,
                                                simulate_batch=_oracle_batch)
    assert got["native_calls"] == 11 and got["status"] == "passed"
    rows = GOLDEN["product"]
    compared = 0
    for key, value in _leaves({"checks": got["checks"], "measurements": got["measurements"]}):
        published = rows[key]["published"]
        if isinstance(published, float) and not isinstance(published, bool):
            assert abs(value - published) <= (1e-6 if published < -120.0 else 1e-11), (key, value, published)
        else:
            assert value == published, (key, value, published)
        compared += 1
    assert compared >= 40
    assert got["measurements"]["roundtrip"]["roundtrip_snr_db"] == rows["/measurements/roundtrip/roundtrip_snr_db"]["published"]


def test_batched_study_fails_the_legacy_configuration_on_the_published_gates():
    got = resampler_eval.evaluate_configuration("legacy-blackman-harris-squared-128", 128, "blackman_harris_squared",
                                                duration_seconds=10, simulate_batch=_oracle_batch)
    failed = sorted(name for name, ok in got["checks"].items() if not ok)
    assert got["status"] == "failed"
    assert failed == ["offline_reference_magnitude", "passband_absolute_error", "passband_ripple", "roundtrip"]  # the report's reason
    rows = GOLDEN["legacy-blackman-harris-squared-128"]
    assert abs(got["measurements"]["passband_and_offline_reference"][0]["max_absolute_error_db"]
               - rows["/measurements/passband_and_offline_reference[0]/max_absolute_error_db"]["published"]) < 1e-11


import pytest  # noqa: E402


@pytest.mark.gpu
def test_batched_study_on_the_gpu_reproduces_the_published_product_configuration():
    got = resampler_eval.evaluate_configuration("product", 128, "blackman", native_default=True, duration_seconds=60)
    assert got["native_calls"] == 11 and got["status"] == "passed"
    rows = GOLDEN["product"]
    for key, value in _leaves({"checks": got["checks"], "measurements": got["measurements"]}):
        published = rows[key]["published"]
        if isinstance(published, float) and not isinstance(published, bool):
            assert abs(value - published) <= (1e-6 if published < -120.0 else 1e-9), (key, value, published)
        else:
            assert value == published, (key, value, published)
