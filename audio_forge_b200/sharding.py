"""Multi-GPU plumbing of a sweep: shard the streams, gather the per-stream metric structs.

Every (candidate, passage) stream is independent end to end, so ranks never exchange audio: each rank
renders its shard and only the ``AfChainMetrics`` structs (136 bytes per stream) are all-gathered --
over NCCL / NVLink on GPUs, over gloo in the CPU tests -- for the final first-safe-scale / argmin pick.
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Sequence

import numpy as np

from . import abi


def shard_streams(costs: Sequence[float], world_size: int) -> list[np.ndarray]:
    """Greedy longest-processing-time split of streams over ranks, balanced by cost (e.g. EQ sections x
    samples: a 48 dB/oct candidate costs up to 4x in the EQ stage).  Deterministic; every rank computes
    the same assignment.  Returns the sorted stream indices of each rank."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    loads = np.zeros(world_size)
    counts = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(costs.size, dtype=np.int64)
    for i in order:
        r = int(np.lexsort((counts, loads))[0])  # least loaded, then fewest streams
        owner[i] = r
        loads[r] += costs[i]
        counts[r] += 1
    return [np.sort(np.nonzero(owner == r)[0]) for r in range(world_size)]


def stream_costs(candidates, pair_candidate: Sequence[int], pair_len: Sequence[int]) -> np.ndarray:
    """Relative render cost of each stream: (fixed chain cost + EQ sections) x samples."""
    sections = []
    for c in candidates:
        n = 0
        for b in c.bands:
            if not c.settings.use_typed_bands:
                n += 1
            elif b.enabled:
                n += b.slope_db_per_octave // 12 if b.filter_type in (4, 5) else 1
        sections.append(n)
    sections = np.asarray(sections, dtype=np.float64)
    return (40.0 + sections[np.asarray(pair_candidate)]) * np.asarray(pair_len, dtype=np.float64)


def metrics_to_bytes(metrics, n: int) -> np.ndarray:
    buf = np.frombuffer(metrics, dtype=np.uint8, count=n * C.sizeof(abi.AfChainMetrics))
    return buf.copy()


def bytes_to_metrics(buf: np.ndarray):
    n = buf.size // C.sizeof(abi.AfChainMetrics)
    arr = (abi.AfChainMetrics * n)()
    C.memmove(arr, np.ascontiguousarray(buf).ctypes.data, n * C.sizeof(abi.AfChainMetrics))
    return arr


def gather_metrics(local_metrics, local_indices: np.ndarray, n_total: int, group=None):
    """All-gather the ranks' metric structs into caller order -> AfChainMetrics[n_total] on every rank.

    ``local_metrics``: AfChainMetrics array (host) of this rank's shard; works with any initialised
    torch.distributed backend (tensors are moved to the GPU for NCCL)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    size = C.sizeof(abi.AfChainMetrics)
    counts = torch.zeros(world, dtype=torch.int64)
    counts[dist.get_rank(group)] = len(local_indices)
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    counts = counts.to(device)
    dist.all_reduce(counts, group=group)
    max_n = int(counts.max().item())
    payload = torch.zeros(max_n * size + max_n * 8, dtype=torch.uint8)
    local = metrics_to_bytes(local_metrics, len(local_indices))
    payload[: local.size] = torch.from_numpy(local)
    idx_bytes = np.asarray(local_indices, dtype=np.int64).view(np.uint8)
    payload[max_n * size: max_n * size + idx_bytes.size] = torch.from_numpy(idx_bytes.copy())
    payload = payload.to(device)
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    out = np.zeros(n_total * size, dtype=np.uint8)
    for r in range(world):
        n = int(counts[r].item())
        blob = gathered[r].cpu().numpy()
        idx = blob[max_n * size: max_n * size + n * 8].view(np.int64)
        rows = blob[: n * size].reshape(n, size)
        out.reshape(n_total, size)[idx] = rows
    return bytes_to_metrics(out)


def sharded_chain_sweep(render, passages, sample_rate: float, candidates, pair_passage, pair_candidate, group=None):
    """The whole multi-GPU step of a sweep: balance the streams over the ranks of ``group``, render this rank's shard
    with ``render(passages, sample_rate, candidates, pair_passage, pair_candidate) -> AfChainMetrics array``
    (``Simulator.chain_sweep``'s first result on a GPU rank), all-gather the metric structs.  Every rank returns the
    AfChainMetrics of ALL pairs in caller order, so the final first-safe-scale / argmin pick is local and identical
    on every rank."""
    import torch.distributed as dist

    pair_passage = np.asarray(pair_passage, dtype=np.uint32)
    pair_candidate = np.asarray(pair_candidate, dtype=np.uint32)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    costs = stream_costs(candidates, pair_candidate, [len(passages[int(p)]) for p in pair_passage])
    mine = shard_streams(costs, world)[rank]
    local = render(passages, sample_rate, candidates, pair_passage[mine], pair_candidate[mine])
    return gather_metrics(local, mine, pair_passage.size, group=group)
