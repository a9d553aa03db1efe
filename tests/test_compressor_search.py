"""Compressor calibration search: the batched search must pick what the reference's sequential search picks.

tests/golden/compressor_search.json was produced by the REFERENCE's own
``voice_setup._calibrate_compressor_threshold`` (pure Python, imported from the reference tree in the build
container) with the CPU oracle as its native door -- see tools/gen_compressor_search_golden.py.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from audio_forge_b200 import abi, compressor_search, headroom, mic_eq_core
from oracle import pyoracle
from tests.signals import speech_like

GOLDEN = json.loads((Path(__file__).parent / "golden" / "compressor_search.json").read_text())


def _oracle_batch(passages, fs, jobs):
    out = []
    for bands, settings in jobs:
        st, _, _ = mic_eq_core.settings_from_mapping(settings)
        m, _, _ = pyoracle.chain_render(passages[0], fs, abi.legacy_bands(bands), st)
        out.append(abi.metrics_to_dict(m))
    return out


def _run(entry, simulate_batch):
    case = entry["case"]
    audio = speech_like(int(case["seconds"] * 48000), seed=case["seed"], level=case["level"])
    return compressor_search.calibrate_compressor_batch(
        speech_audio=audio, sample_rate=48000, eq_settings=GOLDEN["eq_settings"],
        deesser_settings=GOLDEN["deesser_settings"], compressor_settings=case["compressor"],
        simulate_batch=simulate_batch, **case["targets"])


@pytest.mark.parametrize("entry", GOLDEN["cases"], ids=[e["case"]["name"] for e in GOLDEN["cases"]])
def test_batched_search_reproduces_the_reference_decisions_exactly(entry):
    """Oracle renders on both sides -> selected settings, iteration count and objectives are float-identical."""
    calibrated, diag = _run(entry, _oracle_batch)
    for key, want in entry["selected"].items():
        assert calibrated[key] == want, key
    assert diag["iterations"] == entry["iterations"]
    assert diag["expanded_search_selected"] == entry["expanded_search_selected"]
    for key in ("total_objective", "threshold_only_objective", "expanded_candidate_objective", "incumbent_objective"):
        assert diag[key] == entry[key], key
    assert diag["native_calls"] == 3  # phase 1, phase 2, winner verification (the reference makes up to 68)
    margins = diag["decision_margins"]  # the golden decisions are not near-ties: GPU renders may differ by 1e-6 and still agree
    assert margins["candidates"] == entry["iterations"] - 1 and not margins["near_tie"]
    assert margins["best_score_gap"] > compressor_search.TIE_MARGIN


def test_helpers_match_the_reference_definitions():
    assert compressor_search.huber(0.5) == 0.125 and compressor_search.huber(-3.0) == 2.5
    assert compressor_search.halton(1, 2) == 0.5 and abs(compressor_search.halton(5, 3) - (2 / 3 + 1 / 9)) < 1e-15
    assert compressor_search.key_for({"threshold_db": -20.00000049, "ratio": 4, "attack_ms": 10, "release_ms": 200}) == \
        (-20.0, 4.0, 10.0, 200.0)


@pytest.mark.gpu
@pytest.mark.parametrize("entry", GOLDEN["cases"], ids=[e["case"]["name"] for e in GOLDEN["cases"]])
def test_gpu_search_selects_the_same_candidate(entry):
    """GPU renders: the selected candidate and the expanded / threshold-only decision are identical; objectives
    differ only by the device libm (<= 1e-6)."""
    calibrated, diag = _run(entry, mic_eq_core.simulate_auto_eq_chain_batch)
    for key, want in entry["selected"].items():
        assert calibrated[key] == want, key
    assert diag["iterations"] == entry["iterations"]
    assert diag["expanded_search_selected"] == entry["expanded_search_selected"]
    assert abs(diag["total_objective"] - entry["total_objective"]) < 1e-5
