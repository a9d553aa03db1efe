#!/usr/bin/env python
"""bench.py -- chain-simulator throughput on B200 (BASELINE.json metric), one JSON line on stdout.

Default workload (`--workload c5`): the north-star shape, BASELINE config 5 -- the full chain (adaptive hum /
harmonic-notch input cleanup -> dynamic-EQ de-esser -> typed 10-band EQ -> compressor -> lookahead limiter ->
4x true-peak limiter -> detector + fused score reductions) over 8192 candidates x 8 passages x 10 s = 65536
candidate x passage streams.  A "step" is one pass of that whole sweep.  `--gpus N` is STRONG scaling: the same
65536 streams are partitioned over the N ranks by the product's partitioner (`sharding.plan_shards`), every rank
renders its shard through `Simulator` and the per-stream metric structs are all-gathered over NCCL straight from the
sweeps' device tables (`sharding.DeviceGather`) for the final first-safe-scale / argmin pick.  N = 1 renders all
65536 streams on one GPU.

  value      Msamples/s (stream-samples of ALL ranks / max-over-ranks device time), inputs resident in HBM
  e2e        same metric through the public API with HOST buffers (`sharding.sharded_chain_sweep` at N > 1,
             `Simulator.chain_sweep` at N = 1): plan + H2D + render + gather + D2H every step
  parity     the CPU oracle's metrics for a sample of the rendered streams at FULL length against the GPU's:
             mismatches at 0.01 dB / exact counts, and decisions (headroom-safe flag / hard reject) that differ
  roofline   dominant stage kernel: algorithmic bytes of ONE launch (the streams that launch really processes) / its
             mean launch duration vs the measured HBM peak; the binding resource (instruction issue) beside it
  stages     per stage kernel: share, mean launch duration, streams per launch, bound label
  cpu_baseline  the CPU oracle port (the reference's Rust simulator cannot be built here) on all host threads, on a
             bounded sample of the same workload, with the per-core figure
  other_configs  (N = 1 only) BASELINE configs 2, 3 and 4 at full size, each with value / e2e / roofline / parity, and
             `resampler`: the product resampler simulator on its study's 60 s case, 32 signals per call

`--impl reference` times that CPU port alone (rank 0 only).  `--workload c1..c5` picks another BASELINE shape as the
main record (optionally scaled with --candidates / --passages / --seconds).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # one hardware queue per stage stream (before CUDA init)

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from audio_forge_b200 import abi, sharding, workloads  # noqa: E402

METRIC = "chain-sim throughput (candidate x passage stream-samples rendered per second)"
UNIT = "Msamples/s"
FS = workloads.FS

# BASELINE.json configs.  c5 is the north-star target (and the default); the others are sub-records at N = 1.
WORKLOADS = {
    "c1": dict(kind="preset", candidates=1, passages=1, seconds=10.0, level=0.5,
               name="C1 default preset chain render, single stream", chain="DC block + 80 Hz HP -> flat EQ -> compressor -> limiter -> true-peak"),
    "c2": dict(kind="headroom", candidates=4096, passages=1, seconds=30.0, level=0.5,
               name="C2 auto-eq headroom validation", chain="typed EQ -> compressor -> limiter -> true-peak"),
    "c3": dict(kind="compressor_grid", candidates=16384, passages=8, seconds=20.0, level=0.6,
               name="C3 compressor calibration grid", chain="EQ -> compressor (adaptive release) -> limiter -> true-peak"),
    "c4": dict(kind="true_peak", candidates=1, passages=8192, seconds=60.0, level=1.05,
               name="C4 batch true-peak detection + lookahead limiting", chain="limiter -> true-peak limiter -> detector"),
    "c5": dict(kind="full_chain", candidates=8192, passages=8, seconds=10.0, level=0.6,
               name="C5 full chain with de-esser", chain="hum/harmonic notch cleanup -> de-esser -> EQ -> compressor -> limiter -> true-peak"),
}

# Algorithmic bytes per stream-sample of each stage kernel: the f32 / f64 hand-off values it
# must read and write.  Multiplied by the streams ONE launch processes (afsim_sweep_batch_info) and the chunk.
STAGE_BYTES = {"input": 8, "eq": 8, "deesser": 8, "compressor": 8, "limiter": 20, "output": 4,
               "comp_r1": 36, "comp_m2": 48, "comp_r3": 32, "comp_m4": 32, "comp_r5": 16, "comp_m6": 16,
               "lim_m": 12, "lim_r": 16, "tp_fir_in": 8, "tp_r": 12, "tp_fir_out": 4, "tail": 8,
               "de_ra": 36, "de_mb": 88, "de_rc": 88, "de_mc2": 128, "de_rc3": 112, "comp_r7": 16, "input_fanout": 8,
               "de_rc1a": 104, "de_mc1b": 104, "de_rc1c": 56}
# What binds each stage kernel: serial recurrences are dependency-latency bound (one warp walks
# 32 streams), FP64 maps are FP64-pipe / issue bound, the FIR maps FP32-FMA issue bound, the fan-out copy HBM bound.
STAGE_BOUND = {"input": "latency", "input_fanout": "hbm", "eq": "fp64_issue", "deesser": "fp64_issue",
               "compressor": "fp64_issue", "limiter": "fp64_issue", "output": "fp32_issue",
               "comp_r1": "latency", "comp_m2": "fp64_issue", "comp_r3": "latency", "comp_m4": "fp64_issue",
               "comp_r5": "latency", "comp_m6": "fp64_issue", "comp_r7": "latency", "lim_m": "fp32_issue",
               "lim_r": "latency", "tp_fir_in": "fp32_issue", "tp_r": "latency", "tp_fir_out": "fp32_issue",
               "tail": "fp32_issue",
               "de_ra": "latency", "de_mb": "fp64_issue", "de_rc": "latency", "de_mc2": "fp64_issue", "de_rc3": "latency",
               "de_rc1a": "latency", "de_mc1b": "fp64_issue", "de_rc1c": "latency"}
NCU_TABLES = ROOT / "profiles" / "ncu_tables.json"  # per (workload, streams per GPU, chunk): ncu --set full per stage kernel


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS) + ["resampler"],
                    help="BASELINE.json config of the main record; c5 (the north-star target, default)")
    ap.add_argument("--candidates", type=int, default=0, help="override the workload's candidate count")
    ap.add_argument("--passages", type=int, default=0, help="override the workload's passage count")
    ap.add_argument("--seconds", type=float, default=0.0, help="override the workload's passage length")
    ap.add_argument("--cpu-sample-streams", type=int, default=0, help="0 = 4 per host thread")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (also drops `parity`)")
    ap.add_argument("--no-profile", action="store_true", help="skip the serialised per-stage timing pass")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C2 / C3 / C4 sub-records (N = 1)")
    ap.add_argument("--other-configs", default="c2,c3,c4,resampler", help="sub-records to run at N = 1")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._thread = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        def num(v):
            try:
                return float(v)
            except ValueError:
                return None
        sm = [x for x in (num(r[0]) for r in self.rows if r) if x is not None]
        mx = [x for x in (num(r[1]) for r in self.rows if len(r) > 1) if x is not None]
        pw = [x for x in (num(r[2]) for r in self.rows if len(r) > 2) if x is not None]
        reasons = set()
        for r in self.rows:
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(self.rows)}


class DeviceNoise:
    """Stand-in for a list of passages that only exist on the device (hot white noise, workloads.synthetic_noise_host)."""

    def __init__(self, n_passages: int, n_samples: int):
        self.n_passages, self.n_samples = n_passages, n_samples

    def __len__(self):
        return self.n_passages

    def __getitem__(self, i):
        return workloads.synthetic_noise_host(i, self.n_samples)


class Shape:
    """One workload at its (possibly overridden) size: passages, candidates and the candidate-major pair lists.
    The same on every rank and in both arms (seeds do not depend on the rank: the sweep is ONE job)."""

    def __init__(self, name: str, candidates: int = 0, passages: int = 0, seconds: float = 0.0):
        spec = WORKLOADS[name]
        self.key, self.spec = name, spec
        self.n_cand = candidates or spec["candidates"]
        self.n_pass = passages or spec["passages"]
        self.seconds = seconds or spec["seconds"]
        self.n_samples = int(round(self.seconds * FS))
        kind = spec["kind"]
        self.on_device = kind == "true_peak" and self.n_pass * self.n_samples * 4 > (8 << 30)
        if self.on_device:
            # config 4 at full size is 94 GB of input: generated on the device (afsim_sweep_prepare_synthetic); only the
            # few streams the CPU checks are rebuilt on the host from the same counter hash
            self.passages = DeviceNoise(self.n_pass, self.n_samples)
        else:
            self.passages = [workloads.speech_like(self.n_samples, seed=100 + k, level=spec["level"]) for k in range(self.n_pass)]
            if kind == "full_chain":  # config 5: mains hum + harmonic injected at -26 dBFS
                self.passages = [workloads.add_hum(p, 50.37 + 0.11 * k) for k, p in enumerate(self.passages)]
        if kind == "preset":
            self.cands = workloads.default_preset_candidates(self.n_cand)
        elif kind == "headroom":
            self.cands = workloads.headroom_candidates(self.n_cand, seed=1234)
        elif kind == "compressor_grid":
            self.cands = workloads.compressor_grid_candidates(self.n_cand, seed=1234)
        elif kind == "true_peak":
            self.cands = workloads.true_peak_candidates(self.n_cand)
        else:
            self.cands = workloads.full_chain_candidates(self.n_cand, seed=1234)
        self.n_pairs = self.n_cand * self.n_pass
        idx = np.arange(self.n_pairs, dtype=np.int64)
        self.pair_passage = (idx % self.n_pass).astype(np.uint32)  # candidate-major, as afsim_chain_sweep's cross product
        self.pair_candidate = (idx // self.n_pass).astype(np.uint32)

    def config(self, world: int) -> dict:
        return {"workload": self.spec["name"], "candidates": self.n_cand, "passages": self.n_pass,
                "pairs": self.n_pairs, "pairs_per_gpu": (self.n_pairs + world - 1) // world,
                "seconds": self.seconds, "sample_rate": FS, "chain": self.spec["chain"],
                "l2": "inputs larger than L2: the f32/f64 hand-off rings are rewritten every chunk (tens of MB per chunk "
                      "per ring); the shared passages are read through L2 by design",
                "parallelism": (f"{self.n_pairs} candidate x passage streams partitioned over {world} GPUs (sharding.plan_shards), "
                                "NCCL all-gather of the metric structs from device memory" if world > 1 else "single GPU")}


def decision_of(kind: str, m: dict) -> bool:
    """The per-stream decision the callers take from the metrics: headroom-safe (headroom.py:278-289) or, for the
    compressor grid, the search's hard reject (voice_setup.py:862-867 with its -1.5 dB ceiling and 12 dB peak cap)."""
    if kind == "compressor_grid":
        vals = [m[k] for k in abi.METRIC_F32_KEYS]
        return bool((not all(np.isfinite(vals))) or m["output_true_peak_db"] > -1.5 + 0.10
                    or m["compressor_gain_reduction_db"] > 12.0 + 1e-6)
    return bool(workloads.is_headroom_safe(m))


DISCRETE_KEYS = ("true_peak_limited_events", "non_finite_output", "active_analysis_block_count", "processed_samples")


def compare_metrics(kind: str, want, got, picks) -> dict:
    """Oracle metrics of the sampled streams against the GPU's for the same pairs (north_star: 0.01 dB, counts and
    decisions exact)."""
    mism, max_abs, dec = 0, 0.0, 0
    worst = None
    for k, i in enumerate(picks):
        e, g = abi.metrics_to_dict(want[k]), abi.metrics_to_dict(got[int(i)])
        bad = False
        for key in abi.METRIC_F32_KEYS:
            a, b = e[key], g[key]
            if np.isnan(a) and np.isnan(b):
                continue
            if np.isinf(a) or np.isinf(b):
                bad = bad or a != b
                continue
            d = abs(a - b)
            if d > max_abs:
                max_abs, worst = d, key
            bad = bad or d > 0.01
        for key in DISCRETE_KEYS:
            bad = bad or e[key] != g[key]
        mism += int(bad)
        dec += int(decision_of(kind, e) != decision_of(kind, g))
    return {"streams": int(len(picks)), "metric_mismatches": mism, "max_abs_db": max_abs, "worst_key": worst,
            "decision_mismatches": dec, "tolerance_db": 0.01,
            "what": "CPU oracle vs GPU metrics of the sampled streams at full passage length; counts exact; decision = "
                    + ("search hard-reject flag" if kind == "compressor_grid" else "headroom-safe flag")}


def cpu_port_run(shape: Shape, n_sample: int, threads: int):
    """The oracle port on `threads` host threads over n_sample streams spread over the sweep
    -> (Msamples/s, seconds, pair indices, AfChainMetrics of those pairs)."""
    from oracle import pyoracle
    picks = np.unique(np.linspace(0, shape.n_pairs - 1, n_sample).astype(np.int64))
    used = sorted({int(shape.pair_passage[i]) for i in picks})  # materialise only the passages the sample touches
    local = {p: k for k, p in enumerate(used)}
    host = [shape.passages[p] for p in used]
    pp = np.array([local[int(shape.pair_passage[i])] for i in picks], dtype=np.uint32)
    pc = shape.pair_candidate[picks]
    t0 = time.perf_counter()
    metrics = pyoracle.chain_sweep(host, FS, shape.cands, pp, pc, n_threads=threads)
    dt = time.perf_counter() - t0
    return picks.size * shape.n_samples / dt / 1e6, dt, picks, metrics


def run_reference(args, rank: int):
    if rank != 0:
        return
    shape = Shape(args.workload, args.candidates, args.passages, args.seconds)
    threads = os.cpu_count() or 1
    n_sample = args.cpu_sample_streams or min(shape.n_pairs, 4 * threads)
    for _ in range(min(args.warmup, 1)):
        small = Shape(args.workload, min(shape.n_cand, threads), min(shape.n_pass, 2), 1.0)
        cpu_port_run(small, min(small.n_pairs, threads), threads)
    total_s, streams = 0.0, 0
    for _ in range(args.steps):
        _, dt, picks, _ = cpu_port_run(shape, n_sample, threads)
        streams = picks.size
        total_s += dt
    value = args.steps * streams * shape.n_samples / total_s / 1e6
    cores = min(threads, streams)
    sample = f"{streams} of {shape.n_pairs} streams x the full {shape.seconds:g} s passage per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": shape.config(max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "per_core": value / cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "candidates_per_s": value * 1e6 / shape.n_samples,
        "note": "CPU oracle port of the reference's Rust chain simulator (no Rust toolchain in the image), one "
                "stream per host thread; the host has no GPUs to scale over, so the figure is the same at every --gpus",
    }
    emit(line)


def load_ncu_table(key: str, streams: int, chunk: int):
    try:
        tables = json.loads(NCU_TABLES.read_text())
    except Exception:
        return None, None
    name = f"{key}:{streams}:{chunk}"
    return tables.get(name), name


def stage_records(shape: Shape, stages, info, clocks_mhz, fp64_peak, fp32_peak, hbm_peak):
    """Per stage kernel of the first batch: share of the serialised pass, mean launch duration, the streams ONE launch
    processes and what follows from those: algorithmic GB/s, and -- when a committed ncu table of this exact shape
    exists -- instruction-issue and DRAM figures."""
    render = [(i, name, ms, n) for i, (name, ms, n) in enumerate(stages) if name != "finalize"]
    total_ms = sum(ms for _, _, ms, _ in render) or 1.0
    ncu, ncu_key = load_ncu_table(shape.key, info["streams"], info["chunk"])
    issue_peak = 148 * 4 * (clocks_mhz or 1965.0) * 1e6  # warp instructions / s over all SMSPs
    merged: dict = {}
    for i, name, ms, n in render:
        m = merged.setdefault(name, {"ms": 0.0, "launches": 0, "streams": info["stage_streams"][i] if i < len(info["stage_streams"]) else info["streams"]})
        m["ms"] += ms
        m["launches"] += n
    table = []
    for name, m in merged.items():
        launch_ms = m["ms"] / max(m["launches"], 1)
        samples = m["streams"] * info["chunk"]
        row = {"stage": name, "share": m["ms"] / total_ms, "launch_ms": launch_ms, "streams_per_launch": m["streams"],
               "bound": STAGE_BOUND.get(name, "latency"),
               "GBps": STAGE_BYTES.get(name, 8) * samples / (launch_ms * 1e-3) / 1e9 if launch_ms > 0 else None}
        row["hbm_frac"] = row["GBps"] / hbm_peak if row["GBps"] else None
        if row["hbm_frac"] and row["hbm_frac"] > 1.2:  # more than the memory system can move: the hand-off stays in L2
            row["hbm_frac"] = None  # not a fraction of anything: not printed
            row["hbm_note"] = "above the HBM peak: this launch's hand-off rings are L2-resident at this stream count"
        if ncu and name in ncu and launch_ms > 0:
            t = ncu[name]
            row["warp_instr_per_launch"] = t.get("warp_instr")
            row["dram_bytes_per_launch"] = t.get("dram_bytes")
            if t.get("warp_instr"):
                row["issue_frac"] = t["warp_instr"] / (launch_ms * 1e-3) / issue_peak
            for k in ("fp64_pipe_pct", "fma_pipe_pct", "issue_active_pct"):
                if k in t:
                    row[k] = t[k]
        table.append(row)
    return table, ncu_key if ncu else None


def measure(shape: Shape, args, rank: int, world: int, local_rank: int, steps: int, warmup: int, profile: bool,
            cpu_leg: bool, sim=None):
    """One workload through the product on this rank's GPU -> the record (rank 0) or None."""
    import torch
    import torch.distributed as dist

    from audio_forge_b200 import native

    stream = torch.cuda.current_stream()
    own_sim = sim is None
    if own_sim:
        sim = native.Simulator(local_rank, cuda_stream=stream.cuda_stream)
    kind = shape.spec["kind"]
    size = abi.ctypes_sizeof_metrics()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the partition: every rank computes the same one ----------------------------------------------------------
    if world > 1:
        lens = [shape.n_samples] * shape.n_pass
        shards = sharding.plan_shards(shape.cands, shape.pair_passage, shape.pair_candidate, lens, world)
    else:
        shards = [np.arange(shape.n_pairs, dtype=np.int64)]
    mine = shards[rank]
    pp, pc = shape.pair_passage[mine], shape.pair_candidate[mine]

    # ---- resident sweep: value --------------------------------------------------------------------------------------
    if shape.on_device:
        sweep = sim.prepare_synthetic_sweep(1, shape.n_pass, shape.n_samples, FS, shape.cands, pp, pc)
    else:
        sweep = sim.prepare_sweep(shape.passages, FS, shape.cands, pp, pc)
    gather = sharding.DeviceGather(shards, rank) if world > 1 else None

    def step_resident():
        sweep.launch()
        if gather is not None:  # the only collective on the path: the per-stream metric structs, device to device
            gather.gather(sweep.metrics_device_ptr)

    for _ in range(warmup):
        step_resident()
    barrier()
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step_resident()
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = sweep.kernel_count * steps
    metrics = gather.to_host() if gather is not None else sweep.collect()  # all pairs, caller order

    # ---- per-stage timing (serialised pass + live wavefront) and the issue peaks ------------------------------------
    stages, wave, info, fp64_peak, fp32_peak = [], [], None, None, None
    if rank == 0 and profile:
        info = sweep.batch_info()
        stages = sweep.profile_stages(max_chunks=64)
        try:
            wave = sweep.profile_wavefront(first_chunk=64, n_chunks=64)
        except ValueError:
            wave = []
        fp64_peak = sim.issue_peak(0)
        fp32_peak = sim.issue_peak(1)
    sweep.release()
    barrier()

    # ---- end to end through the public API with host buffers ----------------------------------------------------------
    e2e_steps = max(1, min(steps, 3))
    d2h = shape.n_pairs * size  # every rank ends up with the full table
    if shape.on_device:
        # no host copy of a 94 GB batch exists: the end-to-end leg covers generate-on-device + render + D2H metrics
        h2d = shape.n_cand * abi.ctypes_sizeof_candidate_params()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            s2 = sim.prepare_synthetic_sweep(1, shape.n_pass, shape.n_samples, FS, shape.cands, pp, pc)
            s2.launch()
            s2.collect()
            s2.release()
        barrier()
        e2e_s = time.perf_counter() - t0
    else:
        host_passages = [torch.from_numpy(p).pin_memory().numpy() for p in shape.passages]
        h2d = sum(p.nbytes for p in shape.passages) + shape.n_cand * abi.ctypes_sizeof_candidate_params()

        def step_e2e():
            if world > 1:
                return sharding.sharded_chain_sweep(sim, host_passages, FS, shape.cands, shape.pair_passage, shape.pair_candidate)
            return sim.chain_sweep(host_passages, FS, shape.cands, shape.pair_passage, shape.pair_candidate)[0]

        step_e2e()  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        barrier()
        e2e_s = time.perf_counter() - t0

    # ---- reductions over ranks ------------------------------------------------------------------------------------------
    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])
    if own_sim:
        sim.close()
    if rank != 0:
        return None

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    stream_samples = shape.n_pairs * shape.n_samples  # whole job, all ranks
    value = stream_samples * steps / (ms_total * 1e-3) / 1e6
    e2e_value = stream_samples * e2e_steps / e2e_s / 1e6
    clk = clocks.summary()
    roofline, stage_table, wavefront = None, [], None
    if stages and info:
        stage_table, ncu_key = stage_records(shape, stages, info, clk["sm_mhz"], fp64_peak, fp32_peak, hbm_peak)
        dominant = max(stage_table, key=lambda r: r["share"])
        frac = dominant["hbm_frac"]
        roofline = {"bound": "hbm", "achieved": dominant["GBps"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": frac if frac is not None and frac <= 1.2 else None,
                    "traffic": dominant.get("dram_bytes_per_launch"),
                    "kernel": dominant["stage"], "kernel_bound": dominant["bound"], "peak_source": peak_src,
                    "streams_per_launch": dominant["streams_per_launch"], "chunk": info["chunk"],
                    "algorithmic_bytes_per_stream_sample": STAGE_BYTES.get(dominant["stage"], 8),
                    "timing": "mean launch duration of the stage in a serialised pass (CUDA events around every launch "
                              "of 64 chunks); in the timed wavefront the stage kernels overlap",
                    "note": "the chain is FP64 / FP32 issue and dependency-latency bound, not HBM bound (SURVEY 8(d)): the HBM "
                            "fraction is small by construction; `kernel_bound` names what binds the dominant kernel and "
                            "`issue` the measured peaks of that resource",
                    "issue": {"fp64_peak_ginstr_s": fp64_peak, "fp32_fma_peak_ginstr_s": fp32_peak,
                              "unit": "1e9 warp-lane instructions/s, measured in this run",
                              "kernel_issue_frac": dominant.get("issue_frac"), "ncu_table": ncu_key,
                              "wavefront_issue_frac": None}}
        if frac is not None and frac > 1.2:
            roofline["frac_note"] = "not printed: above the HBM peak (L2-resident hand-off at this stream count)"
    if wave and info:
        periods = sorted(p for _, _, p in wave)
        period_ms = periods[len(periods) // 2]
        wavefront = {"period_ms": period_ms, "chunks_timed": 64, "chunk": info["chunk"], "streams": info["streams"],
                     "stages": [{"stage": n, "busy_ms": b} for n, b, _ in wave],
                     "note": "live wavefront (every stage on its own stream): busy_ms = launch eligible -> kernel done, "
                             "under contention with the other stages; period_ms = time between consecutive chunks"}
        ncu, _ = load_ncu_table(shape.key, info["streams"], info["chunk"])
        if ncu:
            instr = sum((ncu.get(n) or {}).get("warp_instr", 0.0) for n, _, _ in wave)
            wavefront["warp_instr_per_chunk"] = instr
            wavefront["issue_frac"] = instr / (period_ms * 1e-3) / (148 * 4 * (clk["sm_mhz"] or 1965.0) * 1e6)
            if roofline:
                roofline["issue"]["wavefront_issue_frac"] = wavefront["issue_frac"]
    record = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": shape.config(world),
        "candidates_per_s": value * 1e6 / shape.n_samples,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "bytes": "per rank"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
        "stages": stage_table,
        "wavefront": wavefront,
    }
    safe = sum(decision_of(kind, abi.metrics_to_dict(metrics[i])) for i in range(shape.n_pairs))
    record["decisions"] = {("hard_rejects" if kind == "compressor_grid" else "safe_streams"): int(safe), "of": shape.n_pairs}
    if cpu_leg:
        threads = os.cpu_count() or 1
        n_sample = args.cpu_sample_streams or min(shape.n_pairs, 4 * threads)
        cpu_value, cpu_s, picks, want = cpu_port_run(shape, n_sample, threads)
        cores = min(threads, int(picks.size))
        record["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": cores, "per_core": cpu_value / cores, "kind": "port",
                                  "sample": f"{picks.size} of {shape.n_pairs} streams x the full passage, {cpu_s:.1f} s"}
        record["parity"] = compare_metrics(kind, want, metrics, picks)
    return record


def measure_resampler(args, local_rank: int, steps: int = 5, warmup: int = 3, n_streams: int = 32, seconds: float = 60.0):
    """the product resampler simulator (`simulate_product_resampler`, resampling.rs:170-262) on the
    long-stream case of the reference's own study (python/tools/evaluate_resampler_quality.py: 60 s, 44.1 -> 48 kHz,
    128-tap Blackman, 256 phases, cubic), `n_streams` signals in one call.  Unit: output frames per second."""
    import torch

    from audio_forge_b200 import native
    from oracle import resampler_oracle  # cpu_baseline / parity legs only

    stream = torch.cuda.current_stream()
    sim = native.Simulator(local_rank, cuda_stream=stream.cuda_stream)
    rate_in, rate_out = 44100, 48000
    n_in = int(rate_in * seconds)
    spec = native.resampler_spec(rate_in, rate_out)
    shape = native.resampler_shape(spec, n_in)
    frames = int(shape.frames)
    rng = np.random.default_rng(0x5EED)
    host = (0.25 * rng.standard_normal((n_streams, n_in))).astype(np.float64)
    d_in = torch.from_numpy(host).to(f"cuda:{local_rank}")
    d_out = torch.zeros((n_streams, frames), dtype=torch.float64, device=f"cuda:{local_rank}")
    kernel_ms = []
    for i in range(warmup + steps):
        ms = sim.product_resampler_device(spec, d_in.data_ptr(), n_in, n_streams, n_in, d_out.data_ptr(), frames)
        if i >= warmup:
            kernel_ms.append(ms)
    torch.cuda.synchronize()
    ms_step = float(np.mean(kernel_ms))
    total_frames = n_streams * frames
    fp64_peak = sim.issue_peak(0)
    fma = total_frames * 4 * int(spec.sinc_len)
    hbm_peak = load_hbm_peak()
    rec = {
        "value": total_frames / (ms_step * 1e-3) / 1e6, "unit": "Mframes/s (output)", "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
        "config": {"workload": "product resampler simulator, the reference study's long-stream case as one batch", "streams": n_streams,
                   "seconds": seconds, "input_rate": rate_in, "output_rate": rate_out, "sinc_len": int(spec.sinc_len), "window": "blackman",
                   "phases": 256, "interpolation": "cubic", "frames_per_stream": frames,
                   "l2": f"inputs larger than L2: {host.nbytes / 1e6:.0f} MB in, {total_frames * 8 / 1e6:.0f} MB out per step"},
        "gpu_launches": steps,
        "roofline": {"bound": "fp64_issue", "kernel": "k_resample", "achieved": fma / (ms_step * 1e-3) / 1e9, "peak": fp64_peak,
                     "unit": "1e9 warp-lane FP64 instructions/s (512 DFMA per frame; peak measured in this run on dependent DMUL+DADD chains)",
                     "frac": fma / (ms_step * 1e-3) / 1e9 / fp64_peak if fp64_peak else None,
                     "hbm_gb_s": (host.nbytes + total_frames * 8) / (ms_step * 1e-3) / 1e9, "hbm_frac": (host.nbytes + total_frames * 8) / (ms_step * 1e-3) / 1e9 / hbm_peak,
                     "algorithmic_bytes_per_frame": 16, "traffic": None},
    }
    # end to end through the public call with host buffers (H2D + render + D2H inside the timed region)
    e2e_s = []
    out = None
    for i in range(1 + min(steps, 2)):
        t0 = time.perf_counter()
        out, _ = sim.product_resampler(host, spec)
        if i:
            e2e_s.append(time.perf_counter() - t0)
    rec["e2e"] = {"value": total_frames / float(np.mean(e2e_s)) / 1e6, "unit": "Mframes/s (output)", "h2d_bytes_per_step": int(host.nbytes),
                  "d2h_bytes_per_step": int(total_frames * 8), "steps": len(e2e_s)}
    if not args.no_cpu_baseline:
        # the oracle on a bounded sample: the first 4 s of two streams (frames sit where the 60 s render puts them, so
        # the same frames of the GPU's 60 s outputs are the comparison)
        n_cpu = int(rate_in * 4.0)
        keep = int(rate_out * 4.0) - 256
        t0 = time.perf_counter()
        worst = 0.0
        for s in range(2):
            want, _, _, _ = resampler_oracle.simulate_product_resampler(host[s, :n_cpu], rate_in, rate_out)
            worst = max(worst, float(np.max(np.abs(want[:keep] - out[s, :keep]))))
        cpu_s = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": 2 * keep / cpu_s / 1e6, "unit": "Mframes/s (output)", "cores": 1, "kind": "port",
                               "sample": f"2 streams x 4 s through the numpy oracle, {cpu_s:.1f} s",
                               "reference_published": {"value": 1024 * rate_out / rate_in / 66.0e-6 / 1e6, "unit": "Mframes/s (output)",
                                                       "what": "the real crate's median 66.0 us per 1024-frame block, evaluation/resampler-quality-report.json (the reference author's machine, one core)"}}
        rec["parity"] = {"streams": 2, "frames": keep, "max_abs_difference": worst, "tolerance": 1e-12,
                         "mismatches": int(worst > 1e-12), "what": "numpy oracle (pinned on the reference's published report) vs GPU, full-scale 1.0"}
    sim.close()
    return rec


def load_hbm_peak() -> float:
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        return 6552.6


def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream(device=local_rank)  # the library launches on it, so torch's events see the work
    torch.cuda.set_stream(stream)

    if args.workload == "resampler":  # the 8(f).4 sub-record alone (one GPU)
        if rank == 0:
            line = measure_resampler(args, local_rank, steps=args.steps, warmup=args.warmup, seconds=args.seconds or 60.0)
            line.update({"metric": "product resampler simulator throughput (output frames per second)", "n_gpus": 1, "higher_is_better": True,
                         "scaling": "replicas only", "vs_baseline": None, "dtype": "f64", "data": "synthetic"})
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return
    shape = Shape(args.workload, args.candidates, args.passages, args.seconds)
    line = measure(shape, args, rank, world, local_rank, args.steps, args.warmup, profile=not args.no_profile,
                   cpu_leg=not args.no_cpu_baseline)
    if rank == 0 and world == 1 and not args.no_other_configs:
        others = {}
        for key in [k for k in args.other_configs.split(",") if (k in WORKLOADS or k == "resampler") and k != args.workload]:
            t0 = time.perf_counter()
            try:
                if key == "resampler":
                    rec = measure_resampler(args, local_rank)
                    rec["wall_s"] = time.perf_counter() - t0
                    others[key] = rec
                    continue
                rec = measure(Shape(key), args, 0, 1, local_rank, steps=2, warmup=3, profile=not args.no_profile,
                              cpu_leg=not args.no_cpu_baseline)
                for drop in ("metric", "unit", "higher_is_better", "vs_baseline", "data", "dtype", "n_gpus", "wavefront"):
                    rec.pop(drop, None)
                rec["wall_s"] = time.perf_counter() - t0
                others[key] = rec
            except Exception as exc:  # a sub-record must never cost the main line
                others[key] = {"error": f"{type(exc).__name__}: {exc}"}
        line["other_configs"] = others
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The one JSON line of the contract, on the real stdout (libraries such as NCCL print banners on fd 1)."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # anything else written to fd 1 from here on goes to stderr
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
