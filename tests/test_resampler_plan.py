"""Host planner of the product resampler simulator (csrc/afsim_resample.cu behind include/afsim.h, no GPU needed)
against the oracle (oracle/resampler_oracle.py, pinned on the reference's published report): output shape, the frame
list -- input window, phase and cubic abscissa of every frame, bit for bit -- the phase table, the reference's validation
(rust-core/src/audio/processor/resampling.rs:187-221) and the reference's own unit tests for the hook
(rust-core/src/audio/processor/tests.rs:193-262) restated."""
import numpy as np
import pytest

from audio_forge_b200 import abi, native
from oracle import resampler_oracle as R

CASES = [
    # input_rate, output_rate, n_in, chunk_size, sinc_len, window
    (44100, 48000, 66150, 1024, None, None),
    (48000, 44100, 72000, 1024, None, None),
    (44100, 48000, 44100, 1024, 128, "blackman_harris_squared"),
    (48000, 44100, 30000, 1024, 256, "blackman_harris_squared"),
    (44100, 48000, 5000, 300, None, None),
    (44100, 48000, 3000, 64, None, None),    # blocks shorter than the filter: some produce no frame
    (44100, 48000, 700, 1, None, None),
    (16000, 48000, 16000, 1024, None, None),
    (96000, 44100, 48000, 1000, None, None),
    (44100, 48000, 0, 1024, None, None),      # nothing in: the flush alone reaches `delay` frames
    (44100, 48000, 1, 1024, None, None),
    (44100, 48000, 1024, 1024, None, None),   # exactly one full block, no partial block
]


@pytest.mark.parametrize("case", CASES)
def test_shape_and_frame_list_equal_the_oracle(case):
    rate_in, rate_out, n_in, chunk, sinc_len, window = case
    spec = native.resampler_spec(rate_in, rate_out, chunk, sinc_len, window)
    shape, _, base, phase, frac = native.resampler_plan(spec, n_in, with_table=False)
    out, delay, expected, _ = R.simulate_product_resampler(np.zeros(n_in), rate_in, rate_out, chunk, sinc_len, window)
    assert (int(shape.frames), int(shape.expected_frames), int(shape.delay)) == (out.size, expected, delay)
    assert shape.frames >= shape.expected_frames + shape.delay  # resampling.rs:245-259
    o_base, o_sub, o_frac = R.frame_list(n_in, rate_in, rate_out, chunk, spec.sinc_len)
    assert np.array_equal(base, o_base)
    assert np.array_equal(phase, o_sub)
    assert np.array_equal(frac, o_frac)  # the positions are accumulated addition for addition: identical bits


def test_flush_block_without_a_frame_is_the_reference_s_runtime_error():
    # one-sample blocks while downsampling: 0.92 frames per block, so a flush block comes up empty (resampling.rs:251-255)
    with pytest.raises(RuntimeError, match="resampler flush produced no frames"):
        R.simulate_product_resampler(np.zeros(700), 48000, 44100, 1)
    with pytest.raises(ValueError, match="resampler flush produced no frames"):
        native.resampler_shape(native.resampler_spec(48000, 44100, 1), 700)


def test_phase_table_equals_the_oracle_to_libm_rounding():
    for sinc_len, window in ((128, "blackman"), (128, "blackman_harris_squared"), (256, "blackman_harris_squared")):
        for rate_in, rate_out in ((44100, 48000), (48000, 44100)):
            spec = native.resampler_spec(rate_in, rate_out, 1024, sinc_len, window)
            _, table, _, _, _ = native.resampler_plan(spec, 2048)
            cutoff = R.effective_cutoff(R.KNOWN_CUTOFFS[(sinc_len, window)], rate_out / rate_in)
            want = R.make_sincs(sinc_len, cutoff, window)
            # glibc sin / cos against numpy's: a few ulp of the largest entries
            assert np.max(np.abs(table - want)) < 4e-16, (sinc_len, window, rate_in)
            assert abs(table.sum() / 256.0 - 1.0) < 1e-12  # unit DC gain per phase on average


def test_reference_unit_test_delay_count_and_length():
    # tests.rs:193-206 product_resampler_offline_hook_reports_delay_count_and_timings
    shape = native.resampler_shape(native.resampler_spec(44100, 48000), 44100)
    assert shape.expected_frames == 48000
    assert shape.delay == 69  # build_sinc_resampler(44100, 48000, 1024).output_delay(): evaluation report, both hooks
    assert shape.frames >= shape.delay + shape.expected_frames
    assert shape.blocks > 0   # one timing entry per process call


def test_reference_unit_test_rejects_invalid_inputs():
    # tests.rs:222-259 product_resampler_offline_hook_rejects_invalid_inputs (the NaN case needs the samples: GPU tier)
    for args, message in (
        ((0, 48000, 1024, None, None), "sample rates must be positive"),
        ((48000, 0, 1024, None, None), "sample rates must be positive"),
        ((48000, 44100, 0, None, None), "chunk_size must be between 1 and 1024"),
        ((48000, 44100, 1025, None, None), "chunk_size must be between 1 and 1024"),
        ((48000, 44100, 1024, 96, None), "sinc_len must be a power of two between 32 and 2048"),
        ((48000, 44100, 1024, 16, None), "sinc_len must be a power of two between 32 and 2048"),
        ((48000, 44100, 1024, 4096, None), "sinc_len must be a power of two between 32 and 2048"),
        ((48000, 44100, 1024, None, "unknown"), "unsupported resampler window"),
    ):
        with pytest.raises(ValueError, match=message):
            native.resampler_shape(native.resampler_spec(*args), 1)


def test_unpinned_configurations_fail_loudly():
    for sinc_len, window in ((64, "blackman"), (128, "hann"), (256, "blackman"), (128, "blackman_harris")):
        with pytest.raises(native.AfsimError) as info:
            native.resampler_shape(native.resampler_spec(44100, 48000, 1024, sinc_len, window), 100)
        assert info.value.status == abi.AFSIM_UNSUPPORTED and "calculate_cutoff" in info.value.message


def test_default_spec_is_the_product_configuration():
    spec = native.resampler_spec(44100, 48000)
    assert (spec.sinc_len, abi.RESAMPLER_WINDOWS[spec.window], spec.chunk_size) == (128, "blackman", 1024)
    from audio_forge_b200 import mic_eq_core
    assert mic_eq_core.product_resampler_configuration() == R.product_resampler_configuration() == (128, "blackman", "cubic", 256, 1024)
