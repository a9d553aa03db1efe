#!/usr/bin/env python
"""bench.py -- chain-simulator throughput on B200 (BASELINE.json metric), one JSON line on stdout.

Workload (N = 1): BASELINE config 2, Auto-EQ headroom validation -- 4096 candidate 10-band typed EQ
settings (25 % with 48 dB/oct Butterworth pass bands) x one 30 s 48 kHz passage through the full
chain (EQ -> compressor -> lookahead limiter -> 4x true-peak limiter -> true-peak detector + fused
score reductions).  A "step" is one pass of that sweep.  N > 1: every rank renders its own 4096
candidates (weak scaling, candidates x passages sharded, no data-path collective) and the
per-candidate metric structs are all-gathered with NCCL for the final first-safe-scale pick.

  value      Msamples/s (stream-samples), inputs resident in HBM, CUDA events on the launching stream
  e2e        same metric through the public API with HOST buffers (plan + H2D + render + D2H per step)
  roofline   dominant stage kernel: algorithmic bytes / its mean launch duration vs measured HBM peak,
             plus `issue`: its FP64 warp-lane instruction rate vs the DMUL+DADD issue peak measured here
  cpu_baseline  the CPU oracle port (the reference's Rust simulator cannot be built here) on all host
             cores, on a bounded sample of the same workload

`--impl reference` times that CPU port alone (rank 0 only).
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

from audio_forge_b200 import abi, workloads  # noqa: E402

METRIC = "chain-sim throughput (candidate x passage stream-samples rendered per second)"
UNIT = "Msamples/s"
FS = workloads.FS

# Algorithmic work per stream-sample of each stage: bytes = f32 read + f32 write
# of the hand-off buffers; fp64 / fp32 = warp-lane arithmetic instructions of the loop body.
STAGE_BYTES = {"input": 8, "eq": 8, "deesser": 8, "compressor": 8, "limiter": 20, "output": 4,
               "comp_r1": 36, "comp_m2": 48, "comp_r3": 32, "comp_m4": 32, "comp_r5": 16, "comp_m6": 16,
               "lim_m": 12, "lim_r": 16, "tp_fir_in": 8, "tp_r": 12, "tp_fir_out": 4}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--candidates", type=int, default=4096)
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--cpu-sample-candidates", type=int, default=0, help="0 = 4 per host thread")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the serialised per-stage timing pass")
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
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def make_workload(args, rank: int):
    n = int(round(args.seconds * FS))
    passage = workloads.speech_like(n, seed=100 + rank, level=0.5)
    cands = workloads.headroom_candidates(args.candidates, seed=1234 + rank)
    return passage, cands


def cpu_port_run(passage, cands, n_sample: int, threads: int):
    """The oracle port on `threads` host threads over the first n_sample candidates -> (Msamples/s, seconds)."""
    from oracle import pyoracle
    sample = (abi.AfCandidate * n_sample).from_buffer(cands)
    pp = np.zeros(n_sample, dtype=np.uint32)
    pc = np.arange(n_sample, dtype=np.uint32)
    t0 = time.perf_counter()
    pyoracle.chain_sweep([passage], FS, sample, pp, pc, n_threads=threads)
    dt = time.perf_counter() - t0
    return n_sample * passage.size / dt / 1e6, dt


def run_reference(args, rank: int):
    if rank != 0:
        return
    passage, cands = make_workload(args, 0)
    threads = os.cpu_count() or 1
    n_sample = args.cpu_sample_candidates or min(args.candidates, 2 * threads)
    for _ in range(min(args.warmup, 1)):
        cpu_port_run(passage[: int(FS)], cands, min(n_sample, threads), threads)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_port_run(passage, cands, n_sample, threads)
        times.append(dt)
    total = sum(times)
    value = args.steps * n_sample * passage.size / total / 1e6
    sample = f"{n_sample} of {args.candidates} candidates x the full {args.seconds:g} s passage per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 auto-eq headroom validation", "candidates": args.candidates,
                   "passages": 1, "seconds": args.seconds, "sample_rate": FS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "candidates_per_s": value * 1e6 / passage.size,
        "note": "CPU oracle port of the reference's Rust chain simulator (no Rust toolchain in the image), one stream per host thread",
    }
    print(json.dumps(line), flush=True)


def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    from audio_forge_b200 import native

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream(device=local_rank)  # the library launches on it, so torch's events see the work
    torch.cuda.set_stream(stream)
    sim = native.Simulator(local_rank, cuda_stream=stream.cuda_stream)
    passage, cands = make_workload(args, rank)
    n_pairs = args.candidates
    stream_samples = n_pairs * passage.size

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident sweep: value ------------------------------------------------------------------------
    sweep = sim.prepare_sweep([passage], FS, cands)
    metrics_bytes = n_pairs * abi.ctypes_sizeof_metrics()
    gathered = None
    if world > 1:
        class _DevBytes:  # zero-copy view of the sweep's device metrics (AfChainMetrics[n_pairs])
            __cuda_array_interface__ = {"shape": (metrics_bytes,), "typestr": "|u1", "version": 2,
                                        "data": (sweep.metrics_device_ptr, False)}
        local = torch.as_tensor(_DevBytes(), device="cuda")
        gathered = torch.empty(world * metrics_bytes, dtype=torch.uint8, device="cuda")

    def step_resident():
        sweep.launch()
        if world > 1:  # the only collective on the path: gather of the per-candidate metric structs
            dist.all_gather_into_tensor(gathered, local)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    with ClockSampler(local_rank) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = sweep.kernel_count * args.steps
    metrics = sweep.collect()

    # ---- end to end through the public API with host buffers --------------------------------------------
    pinned = torch.from_numpy(passage).pin_memory()
    host_passage = pinned.numpy()
    h2d = passage.nbytes + len(cands) * abi.ctypes_sizeof_candidate_params()
    d2h = metrics_bytes
    sim.chain_sweep([host_passage], FS, cands)  # warm-up
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_metrics, _ = sim.chain_sweep([host_passage], FS, cands)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- per-stage timing (serialised pass) + issue peaks --------------------------------------------------
    stages, fp64_peak, fp32_peak = [], None, None
    if rank == 0 and not args.no_profile:
        stages = sweep.profile_stages(max_chunks=64)
        fp64_peak = sim.issue_peak(0)
        fp32_peak = sim.issue_peak(1)
    sweep.release()

    # ---- reductions over ranks ----------------------------------------------------------------------------
    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        value = world * stream_samples * args.steps / (ms_total * 1e-3) / 1e6
        e2e_value = world * stream_samples * e2e_steps / e2e_s / 1e6
        roofline = None
        stage_table = []
        if stages:
            chunk = int(os.environ.get("AFSIM_CHUNK", "1024"))
            total_ms = sum(ms for _, ms, _ in stages if _ != "finalize") or 1.0
            render = [(name, ms, n) for name, ms, n in stages if name != "finalize"]
            merged = {}
            for name, ms, n in render:
                m = merged.setdefault(name, [0.0, 0])
                m[0] += ms
                m[1] += n
            for name, (ms, n) in merged.items():
                per_launch_ms = ms / max(n, 1)
                bytes_per_launch = STAGE_BYTES.get(name, 8) * n_pairs * chunk
                stage_table.append({"stage": name, "share": ms / total_ms, "launch_ms": per_launch_ms,
                                    "GBps": bytes_per_launch / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else None})
            dominant = max(stage_table, key=lambda r: r["share"])
            roofline = {"bound": "hbm", "achieved": dominant["GBps"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": dominant["GBps"] / hbm_peak if dominant["GBps"] else None, "traffic": None,
                        "kernel": dominant["stage"], "peak_source": peak_src,
                        "note": "recurrence kernels are FP64-issue bound, not HBM bound (SURVEY 8(d)); see `issue`",
                        "issue": {"fp64_peak_ginstr_s": fp64_peak, "fp32_fma_peak_ginstr_s": fp32_peak,
                                  "unit": "1e9 warp-lane instructions/s, measured in this run"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2 auto-eq headroom validation", "candidates_per_gpu": args.candidates, "passages": 1,
                       "seconds": args.seconds, "sample_rate": FS, "chain": "typed EQ -> compressor -> limiter -> true-peak",
                       "l2": "work buffers are rewritten every chunk (ring of chunks); 5.8 MB passage stays L2 resident by design",
                       "parallelism": f"candidates sharded x{world}, NCCL all-gather of metric structs" if world > 1 else "single GPU"},
            "candidates_per_s": value * 1e6 / passage.size,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": roofline,
            "stages": stage_table,
            "decisions": {"safe_candidates": int(sum(workloads.is_headroom_safe(abi.metrics_to_dict(metrics[i]))
                                                     for i in range(n_pairs)))},
        }
        if not args.no_cpu_baseline and world >= 1:
            threads = os.cpu_count() or 1
            n_sample = args.cpu_sample_candidates or min(args.candidates, 4 * threads)
            cpu_value, cpu_s = cpu_port_run(passage, cands, n_sample, threads)
            line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n_sample} of {args.candidates} candidates x the full passage, {cpu_s:.1f} s"}
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def main():
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
