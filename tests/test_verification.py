"""Verification renders of Auto Voice Setup (voice_setup.py:1497-1536) in one sweep."""
import numpy as np
import pytest

from audio_forge_b200 import abi, mic_eq_core, verification
from oracle import pyoracle
from tests.signals import speech_like

FS = 48000
SETUP = {
    "eq_settings": {"band_freqs": list(abi.DEFAULT_FREQUENCIES), "band_gains": [2.0, -1.5, 0.0, 1.0, -3.0, 2.5, 4.0, -2.0, 1.5, 0.5],
                    "band_qs": [0.9, 1.41, 1.41, 2.0, 1.2, 1.41, 3.0, 1.41, 1.0, 0.8]},
    "deesser_settings": {"enabled": True, "auto_enabled": True, "auto_amount": 0.6, "max_reduction_db": 8.0},
    "compressor_settings": {"enabled": True, "threshold_db": -26.0, "ratio": 3.0, "attack_ms": 8.0, "release_ms": 150.0,
                            "makeup_gain_db": 6.0, "adaptive_release": True},
}


def _oracle_batch(passages, fs, jobs, *, return_output_audio=False):
    out = []
    for bands, settings in jobs:
        st, _, _ = mic_eq_core.settings_from_mapping(settings)
        for p in passages:
            m, audio, _ = pyoracle.chain_render(p, fs, abi.legacy_bands(bands), st, return_audio=True)
            d = abi.metrics_to_dict(m)
            if return_output_audio:
                d["output_audio"] = audio
            out.append(d)
    return out


def test_chain_and_gates_follow_the_reference():
    eq, chain = verification.verification_chain({})
    assert eq["band_gains"] == [0.0] * 10 and eq["band_qs"] == [1.41] * 10 and eq["band_freqs"] == list(abi.DEFAULT_FREQUENCIES)
    assert chain["limiter"] == {"enabled": True, "ceiling_db": -1.5, "release_ms": 80.0, "careful_output_enabled": True}
    assert chain["return_output_audio"] is True and chain["deesser"] == {} and chain["compressor"] == {}
    eq, chain = verification.verification_chain(SETUP)
    assert eq == SETUP["eq_settings"] and chain["compressor"] == SETUP["compressor_settings"]
    short = np.zeros(FS * 2, dtype=np.float32)
    assert verification.passage_gate(short, FS)["reasons"] == ["verification passage was too short"]
    clipped = speech_like(FS * 4, seed=1, level=0.5)
    clipped[100] = 1.0
    assert verification.passage_gate(clipped, FS)["reasons"] == ["verification passage was non-finite or clipped"]
    bad = speech_like(FS * 4, seed=1, level=0.5)
    bad[7] = np.nan
    assert verification.passage_gate(bad, FS)["decision"] == "retry"
    assert verification.passage_gate(speech_like(FS * 4, seed=1, level=0.5), FS) is None


def test_pair_render_equals_two_sequential_renders_with_the_oracle_door():
    speech = speech_like(FS * 3 + 77, seed=21, level=0.6)
    noise = (speech_like(FS * 2, seed=22, level=0.02)).astype(np.float32)
    processed, rendered, processed_noise, rendered_noise = verification.render_verification_pair(
        noise, speech, FS, SETUP, simulate_batch=_oracle_batch)
    assert rendered.dtype == np.float32 and rendered.size == speech.size and rendered_noise.size == noise.size
    assert processed["simulation_backend"] == "rust" and processed_noise["safety_authority"] == "authoritative"
    assert "output_audio" not in processed and "output_audio" not in processed_noise
    eq, chain = verification.verification_chain(SETUP)
    flat = verification.flatten_chain_settings(chain)
    st, _, _ = mic_eq_core.settings_from_mapping(flat)
    m, audio, _ = pyoracle.chain_render(speech, FS, abi.legacy_bands(verification.bands_from_settings(eq)), st, return_audio=True)
    assert np.array_equal(audio, rendered)
    assert abi.metrics_to_dict(m)["limiter_gain_reduction_db"] == processed["limiter_gain_reduction_db"]
    assert float(np.max(np.abs(rendered))) <= 10 ** (-1.5 / 20) + 1e-6  # the fixed careful limiter of the verification chain


@pytest.mark.gpu
def test_gpu_pair_render_is_bit_identical_to_two_single_calls():
    speech = speech_like(FS * 3 + 77, seed=21, level=0.6)
    noise = (speech_like(FS * 2, seed=22, level=0.02)).astype(np.float32)
    processed, rendered, processed_noise, rendered_noise = verification.render_verification_pair(noise, speech, FS, SETUP)
    eq, chain = verification.verification_chain(SETUP)
    flat = verification.flatten_chain_settings(chain)
    bands = verification.bands_from_settings(eq)
    for audio_in, sim, out in ((speech, processed, rendered), (noise, processed_noise, rendered_noise)):
        single = mic_eq_core.simulate_auto_eq_chain(audio_in, FS, bands, flat)
        assert np.array_equal(np.asarray(single.pop("output_audio"), dtype=np.float32), out)
        for key, value in single.items():
            if key != "candidate_runtime_ms":
                assert sim[key] == value, key


@pytest.mark.gpu
def test_gpu_pair_render_equals_the_oracle():
    """The GPU's verification renders (speech passage + noise capture, one sweep) against the CPU oracle: audio within
    1e-5 / -100 dBFS, every metric within 0.01 dB, counts exact -- what voice_setup.py:1537-1660 computes its spectral
    checks and decision ladder from."""
    from tests.cases import audio_within_tolerance
    speech = speech_like(FS * 3 + 77, seed=21, level=0.6)
    noise = (speech_like(FS * 2, seed=22, level=0.02)).astype(np.float32)
    processed, rendered, processed_noise, rendered_noise = verification.render_verification_pair(noise, speech, FS, SETUP)
    want = verification.render_verification_pair(noise, speech, FS, SETUP, simulate_batch=_oracle_batch)
    for (sim, audio), (o_sim, o_audio) in (((processed, rendered), want[0:2]), ((processed_noise, rendered_noise), want[2:4])):
        assert audio_within_tolerance(o_audio, audio) <= 0.0
        for key, value in o_sim.items():
            if key in ("candidate_runtime_ms", "simulation_backend", "safety_authority"):
                continue
            if isinstance(value, float):
                assert abs(sim[key] - value) <= 0.01 or (np.isinf(value) and sim[key] == value), key
            else:
                assert sim[key] == value, key
