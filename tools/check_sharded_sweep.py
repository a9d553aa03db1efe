"""torchrun check of sharding.sharded_chain_sweep over NCCL: every rank renders its shard on its GPU, all ranks end up
with all metric structs; rank 0 compares a sample with the CPU oracle and the headroom decisions across ranks."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_forge_b200 import abi, native, sharding, workloads  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests.cases import metric_mismatches  # noqa: E402

FS = 48000.0


def main():
    rank, local_rank = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sim = native.Simulator(local_rank)
    passages = [workloads.speech_like(int(FS * 2), seed=700 + k, level=0.6) for k in range(3)]
    cands = workloads.headroom_candidates(70, seed=5)
    pp = np.array([p for c in range(70) for p in range(3)], dtype=np.uint32)
    pc = np.array([c for c in range(70) for p in range(3)], dtype=np.uint32)

    def render(passages, fs, candidates, pair_passage, pair_candidate):
        return sim.chain_sweep(passages, fs, candidates, pair_passage, pair_candidate)[0]

    full = sharding.sharded_chain_sweep(render, passages, FS, cands, pp, pc)
    safe = np.array([workloads.is_headroom_safe(abi.metrics_to_dict(full[i])) for i in range(pp.size)], dtype=np.int64)
    t = torch.from_numpy(safe).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(t, ref))
    bad = 0
    if rank == 0:
        picks = np.arange(0, pp.size, 17)
        want = pyoracle.chain_sweep(passages, FS, cands, pp[picks], pc[picks], n_threads=8)
        bad = sum(1 for k, i in enumerate(picks) if metric_mismatches(want[k], full[int(i)], tol_db=0.01))
    flags = torch.tensor([0 if same else 1, bad], device="cuda")
    dist.all_reduce(flags)
    if rank == 0:
        print(json.dumps({"world": dist.get_world_size(), "pairs": int(pp.size), "decision_mismatch_ranks": int(flags[0]),
                          "oracle_mismatches": int(flags[1]), "safe_pairs": int(safe.sum())}))
    sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
