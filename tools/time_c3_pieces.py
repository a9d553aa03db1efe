"""Render time of compressor-grid sweeps (C3 shape) at several candidate x passage splits, whole and cut
(AFSIM_SUBBATCH): `python tools/time_c3_pieces.py [seconds]` -> one JSON line per shape."""
import json
import os
import sys
from pathlib import Path

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from audio_forge_b200 import native, workloads  # noqa: E402

FS = workloads.FS


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
    sim = native.Simulator(0)
    n = int(seconds * FS)
    for cands_n, pass_n, sub in ((2048, 8, 16384), (16384, 1, 16384), (16384, 2, 16384), (16384, 2, 1 << 30), (16384, 2, 8192)):
        os.environ["AFSIM_SUBBATCH"] = str(sub)
        passages = [workloads.speech_like(n, seed=300 + k, level=0.6) for k in range(pass_n)]
        cands = workloads.compressor_grid_candidates(cands_n)
        sweep = sim.prepare_sweep(passages, FS, cands)
        ms = []
        for _ in range(3):
            sweep.launch()
            sweep.collect()
            ms.append(sweep.render_ms())
        kernels = sweep.kernel_count
        sweep.release()
        best = min(ms)
        print(json.dumps({"candidates": cands_n, "passages": pass_n, "seconds": seconds, "subbatch": sub, "render_ms": ms,
                          "kernels": kernels, "Msamples_s": cands_n * pass_n * n / best / 1e3}), flush=True)
    sim.close()


if __name__ == "__main__":
    main()
