"""The oracle's auto-makeup path (Compressor + the restated `ebur128` momentary loudness meter) against numbers the
REAL reference published: `rust-core/src/bin/auto_makeup_benchmark.rs` drives
`Compressor::process_block_inplace_with_activity_control` with generated tone / noise blocks and injected VAD / noise
evidence, and `evaluation/compressor-control-report.json` holds what the release build printed (nine decimals).

The benchmark's seven arms are restated here against the oracle's Compressor object (signals: `tone_block` :14-21,
`noise_block` :23-31; evidence :33-58; arms :60-181).  The makeup gain is driven by the momentary loudness of the
third-party `ebur128` crate, so agreement here is what pins the oracle's restatement of that crate's meter.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import pyoracle

FS = 48000.0
BLOCK = 480
SPEECH = (1, 0.96, 1.0, -46.0, 1.0)   # :33-40
NOISE = (1, 0.01, 1.0, -38.0, 1.0)    # :42-49
STALE = (1, 0.96, 0.0, -46.0, 0.0)    # :51-58
NONE = (0, 0.0, 0.0, 0.0, 0.0)
# evaluation/compressor-control-report.json, "metrics"
PUBLISHED = {"noise_baseline": 8.434860229, "noise_candidate": 0.0, "speech_baseline": 7.568681717,
             "speech_candidate": 7.568681717, "silence_maximum": 7.568681717, "silence_relaxation": 7.23669083,
             "stale_fresh": 7.568681717, "stale_gain": 2.315e-06, "pumping_baseline_std": 1.139717251,
             "pumping_candidate_std": 0.179404194, "transition_jump": 0.0014644}


class Tone:
    def __init__(self):
        self.phase = 0.0

    def block(self, amplitude, frequency_hz=187.0):
        step = 2.0 * math.pi * frequency_hz / FS
        out = np.empty(BLOCK, dtype=np.float32)
        a = np.float32(amplitude)
        for i in range(BLOCK):
            out[i] = a * np.float32(math.sin(self.phase))
            self.phase = math.fmod(self.phase + step, 2.0 * math.pi)
        return out


class Noise:
    def __init__(self, seed):
        self.state = seed

    def block(self, amplitude):
        units = np.empty(BLOCK, dtype=np.float32)
        for i in range(BLOCK):
            self.state = (self.state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
            units[i] = np.float32((self.state >> 40) & 0xFFFFFFFF)
        unit = units / np.float32((1 << 24) - 1)
        return ((np.float32(2.0) * unit - np.float32(1.0)) * np.float32(amplitude)).astype(np.float32)


class Comp:
    def __init__(self, noise_reference_reliability=1.0):  # :6-12
        L = pyoracle.lib()
        self.L, self.c = L, L.orc_comp_new(-24.0, 3.0, 10.0, 180.0, 0.0, 6.0, FS)
        L.orc_comp_set_auto_makeup_enabled(self.c, 1)
        L.orc_comp_set_target_lufs(self.c, -18.0)
        L.orc_comp_set_noise_reference_reliability(self.c, noise_reference_reliability)

    def run(self, block, evidence):
        buf = np.ascontiguousarray(block, dtype=np.float32).copy()
        self.L.orc_comp_process_block_with_activity(self.c, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size, *evidence)
        return buf

    @property
    def makeup(self):
        return float(self.L.orc_comp_makeup_gain(self.c))

    def close(self):
        self.L.orc_comp_free(self.c)


@pytest.fixture(scope="module")
def measured():
    out = {}
    for name, candidate in (("noise_baseline", False), ("noise_candidate", True)):  # :60-71
        comp, noise = Comp(), Noise(0x6A09E667F3BCC909)
        for _ in range(800):
            comp.run(noise.block(0.022), NOISE if candidate else NONE)
        out[name] = comp.makeup
        comp.close()
    for name, candidate in (("speech_baseline", False), ("speech_candidate", True)):  # :73-84
        comp, tone = Comp(), Tone()
        for _ in range(800):
            comp.run(tone.block(0.035), SPEECH if candidate else NONE)
        out[name] = comp.makeup
        comp.close()
    comp, tone = Comp(), Tone()  # :86-104
    for _ in range(800):
        comp.run(tone.block(0.035), SPEECH)
    speech_gain, maximum = comp.makeup, 0.0
    for _ in range(500):
        comp.run(np.zeros(BLOCK, dtype=np.float32), NOISE)
        maximum = max(maximum, comp.makeup)
    out["silence_maximum"], out["silence_relaxation"] = maximum, speech_gain - comp.makeup
    comp.close()
    comp, tone = Comp(noise_reference_reliability=0.0), Tone()  # :106-120
    for _ in range(800):
        comp.run(tone.block(0.035), SPEECH)
    out["stale_fresh"] = comp.makeup
    for _ in range(300):
        comp.run(tone.block(0.008), STALE)
    out["stale_gain"] = comp.makeup
    comp.close()
    for name, candidate in (("pumping_baseline_std", False), ("pumping_candidate_std", True)):  # :132-160
        comp, tone, noise, history = Comp(), Tone(), Noise(0xBB67AE8584CAA73B), []
        for cycle in range(20):
            for _ in range(40):
                comp.run(tone.block(0.035), SPEECH if candidate else NONE)
                if cycle >= 4:
                    history.append(comp.makeup)
            for _ in range(40):
                comp.run(noise.block(0.012), NOISE if candidate else NONE)
                if cycle >= 4:
                    history.append(comp.makeup)
        h = np.asarray(history, dtype=np.float64)
        out[name] = float(np.sqrt(np.mean((h - h.mean()) ** 2)))
        comp.close()
    comp, jump, previous_last = Comp(), 0.0, np.float32(0.0)  # :162-181
    for index in range(600):
        block = comp.run(np.full(BLOCK, 0.02, dtype=np.float32), SPEECH if (index // 30) % 2 == 0 else NOISE)
        if index > 0:
            jump = max(jump, float(abs(np.float32(block[0] - previous_last))))
        previous_last = block[-1]
    out["transition_jump"] = jump
    comp.close()
    return out


DEVIATING = ("silence_relaxation", "pumping_candidate_std")


@pytest.mark.parametrize("key", sorted(set(PUBLISHED) - set(DEVIATING)))
def test_benchmark_arm_matches_the_published_release_build(measured, key):
    """Printed with nine decimals (`{:.9}`; the report keeps the click with seven, the stale gain with nine).  The
    RMS-only arms (noise 8.434860229 dB, pumping std 1.139717251 dB) and the converged speech gain 7.568681717 dB are
    `target_lufs - momentary loudness` of the restated meter: K-weighting, the 400 ms window and the block feeding
    rule all have to be right for nine decimals."""
    published = PUBLISHED[key]
    tolerance = 5.1e-8 if key == "transition_jump" else 6e-10
    assert abs(measured[key] - published) <= tolerance, (key, measured[key], published)


def test_the_two_arms_that_deviate_from_the_report_are_understood(measured):
    """Two published numbers are NOT reproduced to nine decimals, and are kept visible here rather than loosened away:

    * silence relaxation 7.236690830 published, 7.234470164 here: exactly ONE 10 ms control block of the 1.5 s silence
      relaxation (469 against 468 relaxing blocks out of 500).  With the benchmark source of this tree (silent blocks
      carry `noise_evidence()`, VAD probability 0.01) the smoothed activity score crosses 0.20 at the 33rd silent
      block in exact arithmetic (0.01 + 0.95 e^(-0.05 k) < 0.2  <=>  k > 32.2); the published value is what a
      probability of 0.0 gives (0.96 e^(-0.05 k) < 0.2  <=>  k > 31.4), reproduced below to nine decimals -- the report
      carries no source hash, so it may predate that constant.
    * pumping std of the candidate arm 0.179404194 published, 0.179403283 here (5e-6 relative)."""
    relax = math.exp(-1.0 / 150.0)  # one 480-sample block of MAKEUP_SILENCE_RELAX_MS = 1500
    speech_gain = measured["silence_maximum"]
    assert abs((speech_gain - measured["silence_relaxation"]) - speech_gain * relax ** 468) < 1e-9
    assert abs((speech_gain - PUBLISHED["silence_relaxation"]) - speech_gain * relax ** 469) < 1e-9
    comp, tone = Comp(), Tone()
    for _ in range(800):
        comp.run(tone.block(0.035), SPEECH)
    for _ in range(500):
        comp.run(np.zeros(BLOCK, dtype=np.float32), (1, 0.0, 1.0, -38.0, 1.0))
    assert abs((speech_gain - comp.makeup) - PUBLISHED["silence_relaxation"]) <= 6e-10
    comp.close()
    assert abs(measured["pumping_candidate_std"] - PUBLISHED["pumping_candidate_std"]) < 1.0e-6
