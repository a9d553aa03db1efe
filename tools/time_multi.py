"""The box-wide C-ABI handle (afsim_multi_*, include/afsim.h) on the north-star shape: ONE process, every GPU named in
the mask, the partition + renders + ncclAllGather inside libafsim.so -- what the Rust host would call.
usage: time_multi.py [--mask 0xff] [--candidates 8192] [--passages 8] [--seconds 10] [--steps 3] [--check 32]
Prints one JSON line: Msamples/s from the slowest GPU's device time (CUDA events: render + gather) and from the wall
clock of the call (host buffers in, metrics out), plus an oracle check of `--check` sampled streams."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from audio_forge_b200 import abi, native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mask", default="0x1")
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--candidates", type=int, default=0)
    ap.add_argument("--passages", type=int, default=0)
    ap.add_argument("--seconds", type=float, default=0.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", type=int, default=32)
    args = ap.parse_args()
    shape = bench.Shape(args.workload, args.candidates, args.passages, args.seconds)
    multi = native.MultiSimulator(int(args.mask, 0))
    multi.chain_sweep(shape.passages, bench.FS, shape.cands, shape.pair_passage, shape.pair_candidate)  # warm-up (pools, NCCL)
    wall, dev = [], []
    metrics = None
    for _ in range(args.steps):
        t0 = time.perf_counter()
        metrics = multi.chain_sweep(shape.passages, bench.FS, shape.cands, shape.pair_passage, shape.pair_candidate)
        wall.append(time.perf_counter() - t0)
        dev.append(multi.last_device_ms * 1e-3)
    samples = shape.n_pairs * shape.n_samples
    line = {"tool": "time_multi", "gpus": multi.n_devices, "mask": args.mask, "config": shape.config(multi.n_devices),
            "device_s": float(np.median(dev)), "wall_s": float(np.median(wall)),
            "msamples_per_s_device": samples / float(np.median(dev)) / 1e6, "msamples_per_s_wall": samples / float(np.median(wall)) / 1e6}
    if args.check:
        threads = os.cpu_count() or 1
        _, _, picks, want = bench.cpu_port_run(shape, args.check, threads)
        line["parity"] = bench.compare_metrics(shape.spec["kind"], want, metrics, picks)
    multi.close()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
