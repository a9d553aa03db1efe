"""Wall time of the single-stream entry points on the GPU next to the CPU oracle (C1: one 10 s stream)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_forge_b200 import abi, native  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests.cases import CASES, FS  # noqa: E402
from tests.signals import speech_like  # noqa: E402


def best(fn, reps=5):
    out = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        out.append((time.perf_counter() - t) * 1e3)
    return min(out)


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    x = speech_like(int(FS * seconds), seed=1, level=0.5)
    sim = native.Simulator(0)
    rows = {}
    preset = abi.make_settings(input_stage=1)  # C1: DC block + 80 Hz HP -> default chain
    rows["chain_default_gpu_ms"] = best(lambda: sim.chain_render(x, FS, abi.default_bands(), preset))
    rows["chain_default_cpu_ms"] = best(lambda: pyoracle.chain_render(x, FS, abi.default_bands(), preset), reps=2)
    for name in ("typed_pass", "typed_worst_40_sections"):
        bands, _ = CASES[name]
        rows[f"eq_{name}_gpu_ms"] = best(lambda: sim.eq_render(x, FS, bands, return_audio=True))
        rows[f"eq_{name}_cpu_ms"] = best(lambda: pyoracle.eq_render(x, FS, bands, return_audio=True), reps=2)
    rows["seconds"] = seconds
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
