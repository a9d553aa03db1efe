"""Caller-level wall times: what the reference's UI workers wait for, through the batched B200 callers and through
the CPU oracle one render at a time (the reference renders one candidate per native call, single-threaded, GIL held)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_forge_b200 import abi, compressor_search, headroom, mic_eq_core, workloads  # noqa: E402
from oracle import pyoracle  # noqa: E402

FS = 48000.0


def oracle_batch(passages, fs, jobs):
    out = []
    for bands, settings in jobs:
        st, _, _ = mic_eq_core.settings_from_mapping(settings)
        m, _, _ = pyoracle.chain_render(passages[0], fs, abi.legacy_bands(bands), st)
        out.append(abi.metrics_to_dict(m))
    return out


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return best * 1e3, out


def main():
    golden = json.loads((Path(__file__).resolve().parents[1] / "tests" / "golden" / "compressor_search.json").read_text())
    rng = np.random.default_rng(7)
    speech30 = workloads.speech_like(int(30 * FS), seed=3, level=0.5)
    eq = {"band_freqs": list(abi.DEFAULT_FREQUENCIES), "band_qs": [1.41] * 10,
          "band_gains": [float(g) for g in rng.uniform(-6.0, 9.0, 10)]}
    rows = {}
    mic_eq_core.simulate_auto_eq_chain_batch([speech30[:48000]], FS, [(headroom.bands_from_settings(eq), {})])  # warm-up
    rows["headroom_validation_one_setting_30s_gpu_ms"], a = timed(lambda: headroom.apply_headroom_validation_batch(speech30, FS, [eq]))
    rows["headroom_validation_one_setting_30s_cpu_ms"], b = timed(
        lambda: headroom.apply_headroom_validation_batch(speech30, FS, [eq], simulate_batch=oracle_batch), reps=1)
    rows["headroom_same_scale"] = a[0]["headroom_gain_scale"] == b[0]["headroom_gain_scale"]
    many = [dict(eq, band_gains=[float(g) for g in rng.uniform(-9.0, 9.0, 10)]) for _ in range(64)]
    rows["headroom_validation_64_settings_30s_gpu_ms"], _ = timed(lambda: headroom.apply_headroom_validation_batch(speech30, FS, many), reps=2)
    case = golden["cases"][0]["case"]
    speech20 = workloads.speech_like(int(20 * FS), seed=case["seed"], level=case["level"])

    def search(batch):
        return compressor_search.calibrate_compressor_batch(
            speech_audio=speech20, sample_rate=48000, eq_settings=golden["eq_settings"], deesser_settings=golden["deesser_settings"],
            compressor_settings=case["compressor"], simulate_batch=batch, **case["targets"])

    rows["compressor_search_20s_gpu_ms"], g = timed(lambda: search(mic_eq_core.simulate_auto_eq_chain_batch), reps=2)
    rows["compressor_search_20s_cpu_ms"], c = timed(lambda: search(oracle_batch), reps=1)
    rows["compressor_search_same_selection"] = all(g[0][k] == c[0][k] for k in ("threshold_db", "ratio", "attack_ms", "release_ms"))
    rows["compressor_search_native_calls"] = g[1]["native_calls"]
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
