#!/usr/bin/env python
"""bench.py -- chain-simulator throughput on B200 (BASELINE.json metric), one JSON line on stdout.

Workload at N = 1 (default `--workload c2`): BASELINE config 2, Auto-EQ headroom validation -- 4096
candidate 10-band typed EQ settings (25 % with 48 dB/oct Butterworth pass bands) x one 30 s 48 kHz
passage through the full chain (EQ -> compressor -> lookahead limiter -> 4x true-peak limiter ->
true-peak detector + fused score reductions).  A "step" is one pass of that sweep.  N > 1: every rank
renders its own 4096 candidates (weak scaling; candidates x passages are sharded with no data-path
collective) and the per-candidate metric structs are all-gathered with NCCL for the final
first-safe-scale pick.

  value      Msamples/s (stream-samples), inputs resident in HBM, CUDA events on the launching stream
  e2e        same metric through the public API with HOST buffers (plan + H2D + render + D2H per step)
  roofline   dominant stage kernel: algorithmic bytes / its mean launch duration vs the measured HBM peak,
             plus `issue`: the FP64 / FP32 issue peaks measured in this run (the chain is issue bound)
  stages     per stage kernel: share and mean launch duration in a serialised pass (CUDA events around every launch)
  wavefront  the same batch in the LIVE wavefront: per-stage busy time under contention, the pipeline period and the
             fraction of the GPU's instruction-issue capacity the sweep sustains (`issue_frac`)
  cpu_baseline  the CPU oracle port (the reference's Rust simulator cannot be built here) on all host
             threads, on a bounded sample of the same workload

`--impl reference` times that CPU port alone (rank 0 only).  `--workload c3|c4|c5` runs the other
BASELINE shapes (optionally scaled with --candidates / --passages / --seconds).
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

# BASELINE.json configs.  c2 is the one the metric is quoted on at N = 1; the others are parity / scaling shapes.
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

# Algorithmic bytes per stream-sample of each stage kernel: the f32 / f64 hand-off
# values it must read and write.
STAGE_BYTES = {"input": 8, "eq": 8, "deesser": 8, "compressor": 8, "limiter": 20, "output": 4,
               "comp_r1": 36, "comp_m2": 48, "comp_r3": 32, "comp_m4": 32, "comp_r5": 16, "comp_m6": 16,
               "lim_m": 12, "lim_r": 16, "tp_fir_in": 8, "tp_r": 12, "tp_fir_out": 4,
               "de_ra": 36, "de_mb": 88, "de_rc": 88, "de_mc2": 128, "de_rc3": 112, "comp_r7": 16, "input_fanout": 8,
               "de_rc1a": 104, "de_mc1b": 104, "de_rc1c": 56}


# DRAM traffic per launch (MB, dram__bytes_read.sum + dram__bytes_write.sum) of each stage kernel from the committed
# `ncu --set full` capture of the default workload shape (profiles/r01_v11_ncu_kernels.md: 4096 streams, chunk 1024).
NCU_TRAFFIC_MB = {"input": 0.1, "input_fanout": 0.2, "eq": 16.5, "comp_r1": 92.9, "comp_m2": 160.1, "comp_r3": 82.4,
                  "comp_m4": 109.4, "comp_r5": 35.1, "comp_m6": 50.5, "lim_m": 18.5, "lim_r": 51.6, "tp_fir_in": 17.6,
                  "tp_r": 34.4, "tp_fir_out": 17.5}
# warp-level instructions per launch of the same capture (smsp__inst_executed.sum), for the issue-rate fraction
NCU_WARP_INSTR = {"input": 1.75e4, "input_fanout": 1.73e6, "eq": 8.53e6, "comp_r1": 6.42e6, "comp_m2": 4.18e7,
                  "comp_r3": 4.75e6, "comp_m4": 2.95e7, "comp_r5": 3.79e6, "comp_m6": 1.18e7, "lim_m": 9.41e6,
                  "lim_r": 8.34e6, "tp_fir_in": 2.41e7, "tp_r": 7.88e6, "tp_fir_out": 2.25e7}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS),
                    help="BASELINE.json config; c2 (the headline, default) fits one GPU")
    ap.add_argument("--candidates", type=int, default=0, help="override the workload's candidate count")
    ap.add_argument("--passages", type=int, default=0, help="override the workload's passage count")
    ap.add_argument("--seconds", type=float, default=0.0, help="override the workload's passage length")
    ap.add_argument("--cpu-sample-streams", type=int, default=0, help="0 = 4 per host thread")
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


def make_workload(args, rank: int):
    """-> (passages, candidates); the sweep is the full cross product, candidate-major."""
    spec = WORKLOADS[args.workload]
    args.candidates = args.candidates or spec["candidates"]
    args.passages = args.passages or spec["passages"]
    args.seconds = args.seconds or spec["seconds"]
    n = int(round(args.seconds * FS))
    kind = spec["kind"]
    if kind == "true_peak" and args.passages * n * 4 > (8 << 30):
        # config 4 at full size is 94 GB of input: generated on the device (afsim_sweep_prepare_synthetic); only the
        # few streams the CPU checks are rebuilt on the host from the same counter hash
        return DeviceNoise(args.passages, n), workloads.true_peak_candidates(args.candidates)
    passages = [workloads.speech_like(n, seed=100 + 17 * rank + k, level=spec["level"]) for k in range(args.passages)]
    if kind == "full_chain":  # config 5: mains hum + harmonic injected at -26 dBFS
        passages = [workloads.add_hum(p, 50.37 + 0.11 * k) for k, p in enumerate(passages)]
    if kind == "preset":
        cands = workloads.default_preset_candidates(args.candidates)
    elif kind == "headroom":
        cands = workloads.headroom_candidates(args.candidates, seed=1234 + rank)
    elif kind == "compressor_grid":
        cands = workloads.compressor_grid_candidates(args.candidates, seed=1234 + rank)
    elif kind == "true_peak":
        cands = workloads.true_peak_candidates(args.candidates)
    else:
        cands = workloads.full_chain_candidates(args.candidates, seed=1234 + rank)
    return passages, cands


class DeviceNoise:
    """Stand-in for a list of passages that only exist on the device (hot white noise, workloads.synthetic_noise_host)."""

    def __init__(self, n_passages: int, n_samples: int):
        self.n_passages, self.n_samples = n_passages, n_samples

    def __len__(self):
        return self.n_passages

    def __getitem__(self, i):
        return workloads.synthetic_noise_host(i, self.n_samples)


def config_of(args, world: int) -> dict:
    spec = WORKLOADS[args.workload]
    return {"workload": spec["name"], "candidates_per_gpu": args.candidates, "passages": args.passages,
            "seconds": args.seconds, "sample_rate": FS, "chain": spec["chain"],
            "l2": "the f32/f64 hand-off rings are rewritten every chunk and exceed L2 (tens of MB per chunk per "
                  "ring); the shared passage is read through L2 by design",
            "parallelism": (f"candidate x passage streams sharded x{world}, NCCL all-gather of the metric structs"
                            if world > 1 else "single GPU")}


def cpu_port_run(passages, cands, n_sample: int, threads: int):
    """The oracle port on `threads` host threads over n_sample streams spread over the sweep -> (Msamples/s, s, n)."""
    from oracle import pyoracle
    n_pass, n_cand = len(passages), len(cands)
    total = n_pass * n_cand
    picks = np.unique(np.linspace(0, total - 1, n_sample).astype(np.int64))
    used = sorted({int(i % n_pass) for i in picks})  # materialise only the passages the sample touches
    local = {p: k for k, p in enumerate(used)}
    host = [passages[p] for p in used]
    pp = np.array([local[int(i % n_pass)] for i in picks], dtype=np.uint32)
    pc = (picks // n_pass).astype(np.uint32)
    t0 = time.perf_counter()
    pyoracle.chain_sweep(host, FS, cands, pp, pc, n_threads=threads)
    dt = time.perf_counter() - t0
    return picks.size * host[0].size / dt / 1e6, dt, int(picks.size)


def run_reference(args, rank: int):
    if rank != 0:
        return
    passages, cands = make_workload(args, 0)
    threads = os.cpu_count() or 1
    n_sample = args.cpu_sample_streams or min(len(passages) * len(cands), 2 * threads)
    for _ in range(min(args.warmup, 1)):
        cpu_port_run([passages[i][: int(FS)] for i in range(min(len(passages), 4))], cands, min(n_sample, threads), threads)
    total_s, streams = 0.0, 0
    for _ in range(args.steps):
        _, dt, streams = cpu_port_run(passages, cands, n_sample, threads)
        total_s += dt
    n_samples = passages.n_samples if isinstance(passages, DeviceNoise) else passages[0].size
    value = args.steps * streams * n_samples / total_s / 1e6
    sample = f"{streams} of {len(passages) * len(cands)} streams x the full {args.seconds:g} s passage per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(threads, streams), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "candidates_per_s": value * 1e6 / n_samples,
        "note": "CPU oracle port of the reference's Rust chain simulator (no Rust toolchain in the image), one "
                "stream per host thread",
    }
    emit(line)


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
    passages, cands = make_workload(args, rank)
    on_device = isinstance(passages, DeviceNoise)
    n_pairs = len(cands) * len(passages)
    n_samples = passages.n_samples if on_device else passages[0].size
    stream_samples = n_pairs * n_samples

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident sweep: value ------------------------------------------------------------------------
    if on_device:
        sweep = sim.prepare_synthetic_sweep(1, len(passages), n_samples, FS, cands)
    else:
        sweep = sim.prepare_sweep(passages, FS, cands)
    metrics_bytes = n_pairs * abi.ctypes_sizeof_metrics()
    local = gathered = None
    if world > 1:
        class _DevBytes:  # zero-copy view of the sweep's device metrics (AfChainMetrics[n_pairs])
            __cuda_array_interface__ = {"shape": (metrics_bytes,), "typestr": "|u1", "version": 2,
                                        "data": (sweep.metrics_device_ptr, False)}
        local = torch.as_tensor(_DevBytes(), device="cuda")
        gathered = torch.empty(world * metrics_bytes, dtype=torch.uint8, device="cuda")

    def step_resident():
        sweep.launch()
        if world > 1:  # the only collective on the path: gather of the per-stream metric structs
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

    # ---- per-stage timing (serialised pass) + issue peaks --------------------------------------------------
    stages, wave, fp64_peak, fp32_peak = [], [], None, None
    if rank == 0 and not args.no_profile:
        stages = sweep.profile_stages(max_chunks=64)
        try:  # the same batch in the live wavefront: per-stage duration under contention and the pipeline period
            wave = sweep.profile_wavefront(first_chunk=64, n_chunks=64)
        except ValueError:
            wave = []
        fp64_peak = sim.issue_peak(0)
        fp32_peak = sim.issue_peak(1)
    sweep.release()

    # ---- end to end through the public API with host buffers --------------------------------------------
    e2e_steps = max(1, min(args.steps, 3))
    d2h = metrics_bytes
    if on_device:
        # no host copy of a 94 GB batch exists: the end-to-end leg covers generate-on-device + render + D2H metrics
        h2d = len(cands) * abi.ctypes_sizeof_candidate_params()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            s2 = sim.prepare_synthetic_sweep(1, len(passages), n_samples, FS, cands)
            s2.launch()
            s2.collect()
            s2.release()
        barrier()
        e2e_s = time.perf_counter() - t0
    else:
        host_passages = [torch.from_numpy(p).pin_memory().numpy() for p in passages]
        h2d = sum(p.nbytes for p in passages) + len(cands) * abi.ctypes_sizeof_candidate_params()
        sim.chain_sweep(host_passages, FS, cands)  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sim.chain_sweep(host_passages, FS, cands)
        barrier()
        e2e_s = time.perf_counter() - t0

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
        roofline, stage_table = None, []
        if stages:
            chunk = int(os.environ.get("AFSIM_CHUNK", "1024"))
            render = [(name, ms, n) for name, ms, n in stages if name != "finalize"]
            total_ms = sum(ms for _, ms, _ in render) or 1.0
            merged: dict = {}
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
            default_shape = (args.workload == "c2" and n_pairs == 4096 and chunk == 1024)
            traffic = NCU_TRAFFIC_MB.get(dominant["stage"]) if default_shape else None
            if default_shape:  # instruction-issue view of every stage (the bound that actually binds)
                sm_clock = (clocks.summary()["sm_mhz"] or 1965.0) * 1e6
                issue_peak = 148 * 4 * sm_clock  # warp instructions per second, all SMSPs
                for row in stage_table:
                    wi = NCU_WARP_INSTR.get(row["stage"])
                    if wi and row["launch_ms"]:
                        row["warp_instr_per_s"] = wi / (row["launch_ms"] * 1e-3)
                        row["issue_frac"] = row["warp_instr_per_s"] / issue_peak
            roofline = {"bound": "hbm", "achieved": dominant["GBps"], "peak": hbm_peak, "unit": "GB/s",
                        "frac": dominant["GBps"] / hbm_peak if dominant["GBps"] else None,
                        "traffic": traffic * 1e6 if traffic is not None else None,
                        "kernel": dominant["stage"], "peak_source": peak_src,
                        "timing": "mean launch duration of the stage in a serialised pass (CUDA events around every "
                                  "launch of 64 chunks); in the timed wavefront the stage kernels overlap",
                        "note": "the chain is FP64 / FP32 issue bound, not HBM bound (SURVEY 8(d)); see `issue` and profiles/",
                        "issue": {"fp64_peak_ginstr_s": fp64_peak, "fp32_fma_peak_ginstr_s": fp32_peak,
                                  "unit": "1e9 warp-lane instructions/s, measured in this run",
                                  "wavefront_issue_frac": None}}
        wavefront = None
        if wave:
            periods = sorted(p for _, _, p in wave)
            period_ms = periods[len(periods) // 2]
            wavefront = {"period_ms": period_ms, "chunks_timed": 64,
                         "stages": [{"stage": n, "busy_ms": b} for n, b, _ in wave],
                         "note": "live wavefront (every stage on its own stream): busy_ms = launch eligible -> kernel done, "
                                 "under contention with the other stages; period_ms = time between consecutive chunks"}
            if stages and (args.workload == "c2" and n_pairs == 4096 and int(os.environ.get("AFSIM_CHUNK", "1024")) == 1024):
                sm_clock = (clocks.summary()["sm_mhz"] or 1965.0) * 1e6
                instr = sum(NCU_WARP_INSTR.get(n, 0.0) for n, _, _ in wave)
                wavefront["warp_instr_per_chunk"] = instr
                wavefront["issue_frac"] = instr / (period_ms * 1e-3) / (148 * 4 * sm_clock)
                if roofline:
                    roofline["issue"]["wavefront_issue_frac"] = wavefront["issue_frac"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_of(args, world),
            "candidates_per_s": value * 1e6 / n_samples,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": roofline,
            "stages": stage_table,
            "wavefront": wavefront,
        }
        if WORKLOADS[args.workload]["kind"] == "headroom":
            line["decisions"] = {"safe_candidates": int(sum(workloads.is_headroom_safe(abi.metrics_to_dict(metrics[i]))
                                                            for i in range(n_pairs)))}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_sample = args.cpu_sample_streams or min(n_pairs, 4 * threads)
            cpu_value, cpu_s, streams = cpu_port_run(passages, cands, n_sample, threads)
            line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": min(threads, streams), "kind": "port",
                                    "sample": f"{streams} of {n_pairs} streams x the full passage, {cpu_s:.1f} s"}
        emit(line)
    sim.close()
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
