"""The reference's own unit tests of the auto-makeup controller and the momentary loudness wrapper restated against the
oracle: dsp/compressor.rs:1074-1330 (block-size invariance of the activity smoother to 1e-10,
silence / speech / VAD evidence behaviours, the exact 0.1875 reliability cap, invalid evidence, post-compression
target, limiter-feedback cap, manual makeup fixed) and dsp/loudness.rs:164-220 (creation, invalid rate, silence, a
-20 dBFS 1 kHz tone, reset).  Same constructors, inputs, iteration counts and thresholds as the Rust tests.  The two
`integrated_loudness_lufs` tests (loudness.rs:222-254) are not restated: that helper (`measure_integrated_loudness`,
lib.rs:290-298) is not on the chain simulator's path."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle

FS = 48000.0
AUTO_MAKEUP_ACTIVE_MIN = 0.20  # compressor.rs:25


@pytest.fixture(scope="module")
def L():
    return pyoracle.lib()


class Comp:
    def __init__(self, L, *args):
        self.L, self.p = L, L.orc_comp_new(*[float(a) for a in args])

    @classmethod
    def default_voice(cls, L):  # compressor.rs:205-207
        return cls(L, -20.0, 4.0, 10.0, 200.0, 0.0, 6.0, FS)

    def __getattr__(self, name):
        fn = getattr(self.L, "orc_comp_" + name)
        return lambda *a: fn(self.p, *a)

    def block(self, value, n=48000, evidence=None):
        buf = np.full(n, value, dtype=np.float32)
        if evidence is None:
            self.L.orc_comp_process_block(self.p, pyoracle.fptr(buf), n)
        else:
            self.L.orc_comp_process_block_with_activity(self.p, pyoracle.fptr(buf), n, 1, *[float(v) for v in evidence])
        return buf

    def estimate(self, rms_db, evidence):
        out = np.zeros(2, dtype=np.float64)
        self.L.orc_comp_estimate_activity(self.p, float(rms_db), 1, *[float(v) for v in evidence], pyoracle.dptr(out))
        return float(out[0]), float(out[1])

    def __del__(self):
        self.L.orc_comp_free(self.p)


def test_activity_smoothing_is_block_size_invariant(L):  # compressor.rs:1074-1095
    def activity_after_one_second(block):
        c = Comp.default_voice(L)
        c.set_auto_makeup_enabled(1)
        remaining = 48000
        while remaining > 0:
            elapsed = min(remaining, block)
            c.update_auto_makeup_gain(1.0, 1.0, elapsed)
            remaining -= elapsed
        return c.auto_makeup_activity()

    reference = activity_after_one_second(480)
    for block in (1, 48, 240, 960, 4096, 48000):
        assert abs(activity_after_one_second(block) - reference) < 1e-10, block


def test_does_not_rise_during_silence(L):  # :1097-1109
    c = Comp.default_voice(L)
    c.set_auto_makeup_enabled(1)
    c.set_target_lufs(-12.0)
    for _ in range(4):
        c.block(0.0)
    assert c.makeup_gain() < 0.5


def test_follows_speech_like_blocks(L):  # :1111-1123
    c = Comp.default_voice(L)
    c.set_auto_makeup_enabled(1)
    c.set_target_lufs(-12.0)
    for _ in range(10):
        c.block(0.04)
    assert c.makeup_gain() > 0.1


def test_reliable_vad_prevents_loud_noise_from_driving_makeup(L):  # :1125-1145
    c = Comp.default_voice(L)
    c.set_auto_makeup_enabled(1)
    c.set_target_lufs(-12.0)
    c.set_noise_reference_reliability(1.0)
    for _ in range(10):
        c.block(0.08, evidence=(0.01, 1.0, -32.0, 1.0))
    assert c.auto_makeup_activity() < AUTO_MAKEUP_ACTIVE_MIN
    assert c.makeup_gain() < 0.1


def test_reliable_vad_allows_quiet_speech_to_drive_makeup(L):  # :1147-1167
    c = Comp.default_voice(L)
    c.set_auto_makeup_enabled(1)
    c.set_target_lufs(-12.0)
    for _ in range(10):
        c.block(0.003, evidence=(0.92, 1.0, -68.0, 0.8))
    assert c.auto_makeup_activity() > AUTO_MAKEUP_ACTIVE_MIN
    assert c.auto_makeup_activity_reliability() == 1.0
    assert c.makeup_gain() > 0.1


def test_stale_vad_degrades_continuously_to_noise_relative_fallback(L):  # :1169-1205
    c = Comp.default_voice(L)
    fresh = c.estimate(-52.0, (0.9, 1.0, -55.0, 1.0))
    fading = c.estimate(-52.0, (0.9, 0.5, -55.0, 1.0))
    stale = c.estimate(-52.0, (0.9, 0.0, -55.0, 1.0))
    assert fresh[0] > fading[0] > stale[0]
    assert fresh[1] >= fading[1] >= stale[1]


def test_configured_noise_reliability_cannot_elevate_live_evidence(L):  # :1207-1228
    c = Comp.default_voice(L)
    c.set_noise_reference_reliability(1.0)
    activity, reliability = c.estimate(-53.0, (0.0, 0.0, -60.0, 0.0))
    assert reliability == 0.0
    assert activity == L.orc_comp_speech_activity_from_rms_db(-53.0)


def test_configured_noise_reliability_caps_live_evidence(L):  # :1230-1245
    c = Comp.default_voice(L)
    c.set_noise_reference_reliability(0.25)
    _, reliability = c.estimate(-53.0, (0.0, 0.0, -60.0, 1.0))
    assert abs(reliability - 0.1875) < np.finfo(np.float64).eps


def test_invalid_activity_evidence_cannot_poison_state(L):  # :1247-1265
    c = Comp.default_voice(L)
    c.set_auto_makeup_enabled(1)
    c.set_noise_reference_reliability(float("nan"))
    c.block(0.02, evidence=(float("nan"), float("inf"), float("-inf"), float("nan")))
    assert np.isfinite(c.auto_makeup_activity())
    assert np.isfinite(c.auto_makeup_activity_reliability())
    assert np.isfinite(c.makeup_gain())


def test_targets_post_compression_output_level(L):  # :1267-1288
    compressed = Comp(L, -36.0, 20.0, 0.1, 200.0, 0.0, 0.0, FS)
    uncompressed = Comp(L, 0.0, 1.0, 0.1, 200.0, 0.0, 0.0, FS)
    for c in (compressed, uncompressed):
        c.set_auto_makeup_enabled(1)
        c.set_target_lufs(-12.0)
    for _ in range(10):
        compressed.block(0.04)
        uncompressed.block(0.04)
    assert compressed.gain_reduction() > 1.0
    assert compressed.makeup_gain() >= uncompressed.makeup_gain()


def test_caps_against_limiter_feedback(L):  # :1290-1315
    uncapped, capped = Comp.default_voice(L), Comp.default_voice(L)
    for c in (uncapped, capped):
        c.set_auto_makeup_enabled(1)
        c.set_target_lufs(-12.0)
    capped.set_limiter_feedback_gain_reduction_db(5.0)
    for _ in range(12):
        uncapped.block(0.04)
        capped.block(0.04)
    assert capped.makeup_gain() < uncapped.makeup_gain()
    assert capped.makeup_gain() <= 2.5


def test_manual_makeup_stays_fixed_when_auto_makeup_disabled(L):  # :1317-1330
    c = Comp.default_voice(L)
    c.set_makeup_gain(6.0)
    c.set_auto_makeup_enabled(0)
    for _ in range(4):
        c.block(0.04)
    assert abs(c.makeup_gain() - 6.0) < 1e-9


# ---- dsp/loudness.rs:164-220 ------------------------------------------------------------------------------------------

def test_loudness_meter_creation_and_invalid_rate(L):
    m = L.orc_meter_new(48000)
    assert m
    L.orc_meter_free(m)
    assert not L.orc_meter_new(12345)


def test_loudness_meter_silence(L):
    m = L.orc_meter_new(48000)
    x = np.zeros(48000, dtype=np.float32)
    L.orc_meter_process(m, pyoracle.fptr(x), x.size)
    assert L.orc_meter_momentary(m) < -50.0
    L.orc_meter_free(m)


def test_loudness_meter_tone(L):
    m = L.orc_meter_new(48000)
    i = np.arange(48000, dtype=np.float32)
    phase = (np.float32(2.0) * np.float32(np.pi) * np.float32(1000.0)) * (i / np.float32(48000.0))
    x = (np.float32(0.1) * np.sin(phase, dtype=np.float32)).astype(np.float32)
    L.orc_meter_process(m, pyoracle.fptr(x), x.size)
    lufs = L.orc_meter_momentary(m)
    assert -30.0 < lufs < -10.0
    # BS.1770: a 1 kHz sine at -20 dBFS peak measures -23.0 LUFS (-3.01 dB RMS, +0.0 dB K-weighting at 1 kHz within 0.1)
    assert abs(lufs - (-23.7 + 0.691)) < 0.15
    L.orc_meter_free(m)


def test_loudness_meter_reset_restores_idle_state(L):
    m = L.orc_meter_new(48000)
    x = np.full(48000, 0.1, dtype=np.float32)
    L.orc_meter_process(m, pyoracle.fptr(x), x.size)
    L.orc_meter_reset(m)
    assert L.orc_meter_momentary(m) == -100.0
    L.orc_meter_free(m)
