"""Stage profile of a sweep in the live chunk x stage wavefront and in a serialised pass, plus the host wall time
of sweep.launch() (issue only) next to the GPU time of the same launch.

usage: profile_wavefront.py [candidates [seconds [c2|c3|c5 [passages]]]]   (AFSIM_CHUNK / AFSIM_SLOTS are honoured)"""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_forge_b200 import native, workloads  # noqa: E402

FS = 48000.0


def main():
    n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
    kind = sys.argv[3] if len(sys.argv) > 3 else "c2"
    n_pass = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    sim = native.Simulator(0)
    passages = [workloads.speech_like(int(FS * seconds), seed=100 + k, level=0.5) for k in range(n_pass)]
    if kind == "c5":
        passages = [workloads.add_hum(p, 50.37 + 0.11 * k) for k, p in enumerate(passages)]
        cands = workloads.full_chain_candidates(n_cand, seed=1234)
    elif kind == "c3":
        cands = workloads.compressor_grid_candidates(n_cand, seed=1234)
    else:
        cands = workloads.headroom_candidates(n_cand, seed=1234)
    sweep = sim.prepare_sweep(passages, FS, cands)
    out = {"chunk": os.environ.get("AFSIM_CHUNK", "1024"), "slots": os.environ.get("AFSIM_SLOTS", "default"),
           "candidates": n_cand, "kernels": sweep.kernel_count}
    for rep in range(3):
        t0 = time.perf_counter()
        sweep.launch()
        t1 = time.perf_counter()
        gpu_ms = sweep.render_ms()
        t2 = time.perf_counter()
        out[f"rep{rep}"] = {"host_issue_ms": (t1 - t0) * 1e3, "gpu_ms": gpu_ms, "wall_ms": (t2 - t0) * 1e3}
    out["kind"], out["passages"], out["seconds"] = kind, n_pass, seconds
    out["wavefront"] = [(n, round(b, 4), round(p, 4)) for n, b, p in sweep.profile_wavefront(40, 64)]
    out["serial"] = [(n, round(ms / k, 4)) for n, ms, k in sweep.profile_stages(max_chunks=64)]
    sweep.release()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
