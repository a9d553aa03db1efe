"""Diagnostic: ONE rank's shard of the N-way partitioned north-star sweep rendered alone on one GPU, against the proxy
shapes (`bench.py --candidates 1024`: the first 1024 candidates x all passages; a contiguous-candidate partition).
Separates what the shard's composition costs from what eight processes on one box cost.
usage: time_shard.py [--world 8] [--rank 0] [--steps 3]  -> one JSON line"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from audio_forge_b200 import native, sharding  # noqa: E402


def render(sim, shape, picks, steps):
    pp, pc = shape.pair_passage[picks], shape.pair_candidate[picks]
    sweep = sim.prepare_sweep(shape.passages, bench.FS, shape.cands, pp, pc)
    ms = []
    for i in range(1 + steps):
        sweep.launch()
        if i:
            ms.append(sweep.render_ms())
    info = sweep.batch_info()
    sweep.release()
    return {"streams": int(picks.size), "ms": float(np.median(ms)), "batches": info.get("batches"), "stages": info.get("stages"), "stage_streams": info.get("stage_streams")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    shape = bench.Shape("c5", 0, 0, 0.0)
    sim = native.Simulator(0)
    lens = [shape.n_samples] * shape.n_pass
    shards = sharding.plan_shards(shape.cands, shape.pair_passage, shape.pair_candidate, lens, args.world)
    per = shape.n_cand // args.world
    lo = args.rank * per
    contiguous = np.flatnonzero((shape.pair_candidate >= lo) & (shape.pair_candidate < lo + per))
    by_passage = np.flatnonzero(shape.pair_passage == args.rank % shape.n_pass)
    line = {"tool": "time_shard", "world": args.world, "rank": args.rank,
            "lpt_shard": render(sim, shape, shards[args.rank], args.steps),
            "contiguous_candidates": render(sim, shape, contiguous, args.steps),
            "one_passage_all_candidates": render(sim, shape, by_passage, args.steps)}
    sim.close()
    print(json.dumps(line))


if __name__ == "__main__":
    main()
