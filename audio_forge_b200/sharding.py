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
    import heapq

    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    heap = [(0.0, 0, r) for r in range(world_size)]  # (load, streams, rank): least loaded, then fewest streams
    owner = np.empty(costs.size, dtype=np.int64)
    for i in order.tolist():
        load, count, r = heapq.heappop(heap)
        owner[i] = r
        heapq.heappush(heap, (load + float(costs[i]), count + 1, r))
    return [np.sort(np.nonzero(owner == r)[0]) for r in range(world_size)]


def shard_units(costs: Sequence[float], group_key: Sequence[int], world_size: int) -> list[np.ndarray]:
    """The partition rule of a sweep: streams that share a candidate stay on one rank.

    A rank that holds ONE passage of every candidate reads every candidate's constants (2.7 KB each, per stage kernel and
    chunk) for a single stream; a rank that holds ALL passages of an eighth of the candidates reads an eighth of them,
    each shared by its passages' streams (north-star shape on 8 GPUs: 601 against 662 ms per sweep).  So the units of the
    longest-processing-time split are the streams of one candidate, in caller order, cut into pieces of at most
    ceil(n / (16 world)) streams so that a sweep of few candidates over many passages still spreads over all ranks."""
    import heapq

    costs = np.asarray(costs, dtype=np.float64)
    keys = np.asarray(group_key, dtype=np.int64)
    cap = max(1, -(-costs.size // (16 * world_size)))
    unit_of = np.empty(costs.size, dtype=np.int64)
    open_unit: dict[int, int] = {}
    unit_cost: list[float] = []
    unit_size: list[int] = []
    for i, k in enumerate(keys.tolist()):
        u = open_unit.get(k, -1)
        if u < 0 or unit_size[u] >= cap:
            u = len(unit_cost)
            open_unit[k] = u
            unit_cost.append(0.0)
            unit_size.append(0)
        unit_of[i] = u
        unit_cost[u] += float(costs[i])
        unit_size[u] += 1
    ucost = np.asarray(unit_cost, dtype=np.float64)
    order = np.argsort(-ucost, kind="stable")
    heap = [(0.0, 0, r) for r in range(world_size)]  # (load, streams, rank): least loaded, then fewest streams
    unit_owner = np.empty(ucost.size, dtype=np.int64)
    for u in order.tolist():
        load, count, r = heapq.heappop(heap)
        unit_owner[u] = r
        heapq.heappush(heap, (load + float(ucost[u]), count + unit_size[u], r))
    owner = unit_owner[unit_of] if costs.size else np.empty(0, dtype=np.int64)
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


class _DeviceBytes:
    """Zero-copy view of device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "version": 2, "data": (int(ptr), False)}


class DeviceGather:
    """The all-gather of a sharded sweep's metric structs WITHOUT a host hop: the sweep's device table
    (``afsim_sweep_metrics_device_ptr``) is the NCCL send buffer, the gathered table is permuted into caller order on
    the device and leaves the GPU once.  Buffers are built once per (shard layout) and reused by every step."""

    def __init__(self, shards: Sequence[np.ndarray], rank: int, group=None):
        import torch

        self.size = C.sizeof(abi.AfChainMetrics)
        self.group, self.rank, self.world = group, rank, len(shards)
        self.counts = [int(len(s)) for s in shards]
        self.n_total = int(sum(self.counts))
        self.max_n = max(self.counts) if self.counts else 0
        dev = torch.device("cuda", torch.cuda.current_device())
        self.payload = torch.zeros(max(self.max_n, 1) * self.size, dtype=torch.uint8, device=dev)
        self.gathered = torch.zeros(self.world * max(self.max_n, 1) * self.size, dtype=torch.uint8, device=dev)
        # row of the gathered [world * max_n] table that holds caller pair i
        src = np.zeros(self.n_total, dtype=np.int64)
        for r, idx in enumerate(shards):
            src[np.asarray(idx, dtype=np.int64)] = r * self.max_n + np.arange(len(idx), dtype=np.int64)
        self.src_rows = torch.from_numpy(src).to(dev)
        self.ordered = torch.zeros(max(self.n_total, 1) * self.size, dtype=torch.uint8, device=dev)

    def gather(self, metrics_device_ptr: int):
        """Queue copy + all-gather + permutation on torch's current stream -> device tensor [n_total * 136] (caller order)."""
        import torch
        import torch.distributed as dist

        n_local = self.counts[self.rank]
        if n_local:
            local = torch.as_tensor(_DeviceBytes(metrics_device_ptr, n_local * self.size), device=self.payload.device)
            self.payload[: n_local * self.size].copy_(local, non_blocking=True)
        dist.all_gather_into_tensor(self.gathered, self.payload, group=self.group)
        if self.n_total:
            rows = self.gathered.view(self.world * max(self.max_n, 1), self.size)
            torch.index_select(rows, 0, self.src_rows, out=self.ordered.view(max(self.n_total, 1), self.size)[: self.n_total])
        return self.ordered

    def to_host(self):
        """AfChainMetrics[n_total] on the host (one D2H copy; synchronises)."""
        return bytes_to_metrics(self.ordered[: self.n_total * self.size].cpu().numpy())


def plan_shards(candidates, pair_passage, pair_candidate, passage_lens, world: int) -> list[np.ndarray]:
    """The partition every rank computes identically: the streams of a candidate stay together (`shard_units`), ranks
    balanced by (fixed chain cost + EQ sections) x samples."""
    pair_passage = np.asarray(pair_passage, dtype=np.int64)
    costs = stream_costs(candidates, pair_candidate, np.asarray(passage_lens, dtype=np.float64)[pair_passage])
    return shard_units(costs, pair_candidate, world)


def sharded_chain_sweep(render, passages, sample_rate: float, candidates, pair_passage, pair_candidate, group=None):
    """The whole multi-GPU step of a sweep: balance the streams over the ranks of ``group``, render this rank's shard,
    all-gather the metric structs.  Every rank returns the AfChainMetrics of ALL pairs in caller order, so the final
    first-safe-scale / argmin pick is local and identical on every rank.

    ``render`` is a ``native.Simulator`` (GPU ranks, NCCL group: the shard's metrics are gathered straight from the
    sweep's device table, no host hop) or any callable ``render(passages, sample_rate, candidates, pair_passage,
    pair_candidate) -> AfChainMetrics array`` (the gloo tests on CPU)."""
    import torch.distributed as dist

    pair_passage = np.asarray(pair_passage, dtype=np.uint32)
    pair_candidate = np.asarray(pair_candidate, dtype=np.uint32)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    shards = plan_shards(candidates, pair_passage, pair_candidate, [len(p) for p in passages], world)
    mine = shards[rank]
    if hasattr(render, "prepare_sweep") and dist.get_backend(group) == "nccl":
        import torch

        sweep = render.prepare_sweep(passages, sample_rate, candidates, pair_passage[mine], pair_candidate[mine])
        try:
            gather = DeviceGather(shards, rank, group)
            if render.cuda_stream:  # the library launches on the caller's stream: queue the collective behind it
                with torch.cuda.stream(torch.cuda.ExternalStream(render.cuda_stream)):
                    sweep.launch()
                    gather.gather(sweep.metrics_device_ptr)
                    out = gather.to_host()
            else:  # the library's own stream: wait for the render, then gather on torch's current stream
                sweep.launch()
                sweep.render_ms()
                gather.gather(sweep.metrics_device_ptr)
                out = gather.to_host()
        finally:
            sweep.release()
        return out
    if hasattr(render, "chain_sweep"):
        local = render.chain_sweep(passages, sample_rate, candidates, pair_passage[mine], pair_candidate[mine])[0]
    else:
        local = render(passages, sample_rate, candidates, pair_passage[mine], pair_candidate[mine])
    return gather_metrics(local, mine, pair_passage.size, group=group)
