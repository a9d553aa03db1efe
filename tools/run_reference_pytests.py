"""Runs the REFERENCE's own pytest files for the chain-simulator path, unmodified, against this repository's
reference-facing module (`audio_forge_b200/mic_eq_core.py`: argument parsing, validation, error messages, result
dicts) with the CPU oracle standing in for the GPU (`--door oracle`, the default: the build container has no GPU)
or with the real library (`--door gpu`, on a B200 box that also has the reference tree).

`mic_eq/__init__.py:38-46` picks up a top-level `mic_eq_core`; this script registers the product module under that
name, so `from mic_eq import simulate_auto_eq_chain, ...` in the reference's tests resolves to it.  Nothing of the
reference is copied; the files are collected where they lie under /root/reference (PyQt6 is absent, hence
--noconftest).

usage: python tools/run_reference_pytests.py [--door oracle|gpu] [--fast] [extra pytest args]
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

FILES = ["test_eq_filter_types.py", "test_dynamics_aliasing_tools.py", "test_limiter_lookahead_tools.py",
         "test_processing_order_tools.py", "test_resampler_quality_tools.py", "test_auto_makeup_real_speech_tools.py", "test_auto_eq.py", "test_voice_setup.py"]
# Not collected: test_eq_native_response.py imports the Qt curve widget (PyQt6 is not in this image); its two native
# assertions are restated in tests/test_reference_contract.py.
DESELECT = [
    # the live engine's AudioProcessor class (out of scope: the realtime callback stays on the CPU)
    "test_eq_filter_types.py::test_typed_native_eq_api_round_trips_every_runtime_field",
    "test_eq_filter_types.py::test_typed_native_eq_api_rejects_unknown_type_and_odd_slope",
    "test_eq_filter_types.py::test_legacy_batch_api_restores_historical_filter_layout",
    # fails in this image with NO native core as well (pure-Python Auto-EQ optimiser, numpy / scipy versions)
    "test_auto_eq.py::test_05_proximity_effect_correction",
]


class OracleSimulator:
    """The subset of native.Simulator that mic_eq_core.py calls, answered by the oracle."""

    def chain_render(self, audio, sample_rate, bands, settings, return_audio=False):
        from oracle import pyoracle
        m, out, _ = pyoracle.chain_render(audio, sample_rate, bands, settings, return_audio=return_audio)
        return m, out

    def eq_render(self, audio, sample_rate, bands, return_audio=False):
        from oracle import pyoracle
        return pyoracle.eq_render(audio, sample_rate, bands, return_audio=return_audio)

    def auto_makeup_control(self, audio, sample_rate, vad, noise_floor_db, noise_reliability, settings, return_audio=False):
        from oracle import pyoracle
        return pyoracle.auto_makeup_control(audio, sample_rate, vad, noise_floor_db, noise_reliability, settings,
                                            return_audio=return_audio)

    def product_resampler(self, signals, spec):
        from audio_forge_b200 import abi, native
        from oracle import resampler_oracle
        window = abi.RESAMPLER_WINDOWS[spec.window]
        rows = [resampler_oracle.simulate_product_resampler(row, spec.input_rate, spec.output_rate, spec.chunk_size, spec.sinc_len, window)[0]
                for row in np.asarray(signals, dtype=np.float64)]
        return np.stack(rows), native.resampler_shape(spec, np.asarray(signals).shape[1])

    def eq_response(self, freqs, bands, sample_rate, typed):
        from oracle import pyoracle
        return (pyoracle.eq_response(freqs, bands, sample_rate, typed=typed),)


FAST = FILES[:5]  # seconds; the voice-setup and Auto-EQ files render hundreds of passages through the CPU oracle


def main(argv):
    door = "oracle"
    files_to_run = FILES
    if "--fast" in argv:
        argv.remove("--fast")
        files_to_run = FAST
    if "--door" in argv:
        i = argv.index("--door")
        door = argv[i + 1]
        del argv[i:i + 2]
    from audio_forge_b200 import mic_eq_core as product

    if door == "oracle":
        oracle_sim = OracleSimulator()
        product.simulator = lambda device=0: oracle_sim
    sys.modules["mic_eq_core"] = product
    sys.path.insert(0, str(REF / "python"))
    sys.path.insert(0, str(REF / "python" / "tools"))
    import pytest

    tests = REF / "python" / "tests"
    files = [str(tests / f) for f in files_to_run]
    skip = "not (" + " or ".join(d.split("::")[1] for d in DESELECT) + ")"
    return pytest.main(["--noconftest", "-p", "no:cacheprovider", "--rootdir", "/tmp", "-q", "-k", skip, *files, *argv])


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1:]))
