"""Speech-aware auto-makeup scoring (audio_forge_b200/makeup_eval.py) against the reference's own helpers
(golden vectors made by tools/gen_makeup_eval_golden.py from evaluate_auto_makeup_real_speech.py, unmodified) and,
on the GPU, the batched clip scoring against the same statistics over oracle renders."""
import json
from pathlib import Path

import numpy as np
import pytest

from audio_forge_b200 import abi, makeup_eval
from tests.signals import lcg_noise, speech_like

GOLDEN = json.loads((Path(__file__).parent / "golden" / "makeup_eval.json").read_text())


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: f"n{c['n']}")
def test_helpers_match_the_reference(case):
    n = case["n"]
    audio = lcg_noise(n, state=case["lcg_state"], scale=case["lcg_scale"])[0].astype(np.float32)
    blocks = (n + 479) // 480
    assert np.array_equal(makeup_eval.control_probabilities(case["frames"], n, blocks), np.asarray(case["control"]))
    assert np.array_equal(makeup_eval.block_rms_db(audio), np.asarray(case["block_rms_db"]))
    assert makeup_eval.pumping_score(case["trace"]) == case["pumping"]


def test_pumping_score_prefers_two_to_eight_hz_modulation():
    """test_auto_makeup_real_speech_tools.py:65-70 restated."""
    time = np.arange(1_000) / makeup_eval.CONTROL_CADENCE_HZ
    fast, slow = np.sin(2.0 * np.pi * 4.0 * time), np.sin(2.0 * np.pi * 0.2 * time)
    assert makeup_eval.pumping_score(fast) > 5.0 * makeup_eval.pumping_score(slow)


@pytest.mark.gpu
def test_batched_clip_scoring_matches_oracle_renders():
    from audio_forge_b200 import native
    from oracle import pyoracle
    sim = native.Simulator(0)
    clips = []
    for k in range(3):
        noisy = speech_like(48000 * 2 + 700 * k, seed=40 + k, level=0.3)
        blocks = (noisy.size + 479) // 480
        clean = np.clip(0.5 + 0.5 * np.sin(np.arange(blocks) * 0.04 + k), 0.0, 1.0)
        clips.append((noisy, clean, np.clip(clean + 0.05, 0.0, 1.0)))
    got = makeup_eval.score_clips(sim, clips)
    st = abi.make_makeup_settings(vad_reliability=1.0, adaptive_release=True)
    for (noisy, clean, noisy_control), row in zip(clips, got):
        floor = row["noise_floor_db"]
        cand_t, cand_a = pyoracle.auto_makeup_control(noisy, 48000.0, noisy_control, floor, 1.0, st, return_audio=True)
        base_t, base_a = pyoracle.auto_makeup_control(noisy, 48000.0, None, floor, 1.0, st, return_audio=True)
        want = makeup_eval.clip_metrics(noisy, clean, cand_t[0], base_t[0], cand_a, base_a)
        for key, value in want.items():
            assert abs(row[key] - value) <= 1e-2 * max(1.0, abs(value)), (key, row[key], value)
    sim.close()
