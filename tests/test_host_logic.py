"""Host-side logic that needs no GPU: C-ABI surface, batched headroom decisions, sharding + gloo gather."""
import ctypes as C
import os
import re
import socket
from pathlib import Path

import numpy as np
import pytest

from audio_forge_b200 import abi, headroom, native, sharding
from oracle import pyoracle
from tests.cases import FS, candidate, candidate_array, metric_mismatches
from tests.signals import speech_like

ROOT = Path(__file__).resolve().parents[1]


def test_library_loads_and_exports_every_declared_symbol():
    """Every function include/afsim.h declares is exported by libafsim.so (no compute calls here)."""
    native.build()
    header = (ROOT / "include" / "afsim.h").read_text()
    declared = set(re.findall(r"\b(afsim_[a-z_0-9]+)\s*\(", header))
    lib = C.CDLL(str(native.LIB_PATH))
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert missing == []
    assert declared == set(native.EXPORTS)
    lib.afsim_abi_version.restype = C.c_int
    assert lib.afsim_abi_version() == 1


def test_pod_layouts_match_the_header():
    assert C.sizeof(abi.AfBand) == 32
    assert C.sizeof(abi.AfChainSettings) == 16 + 18 * 8
    assert C.sizeof(abi.AfCandidate) == 10 * 32 + 160
    assert C.sizeof(abi.AfChainMetrics) == 136
    lib = C.CDLL(str(native.LIB_PATH))
    s = abi.AfChainSettings()
    lib.afsim_chain_settings_default(C.byref(s))
    ref = abi.make_settings()
    assert bytes(s) == bytes(ref)
    bands = (abi.AfBand * 10)()
    lib.afsim_default_bands(bands)
    assert bytes(bands) == bytes(abi.default_bands())


def test_no_gpu_means_a_loud_failure_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(native.AfsimError, match="no CPU path"):
        native.Simulator(0)


def _oracle_batch(passages, fs, jobs):
    """Stand-in for mic_eq_core.simulate_auto_eq_chain_batch built on the CPU oracle (tests only)."""
    from audio_forge_b200 import mic_eq_core
    out = []
    for bands, settings in jobs:
        st, _, _ = mic_eq_core.settings_from_mapping(settings)
        m, _, _ = pyoracle.chain_render(passages[0], fs, abi.legacy_bands(bands), st)
        out.append(abi.metrics_to_dict(m))
    return out


def _sequential_reference_walk(audio, fs, eq_settings, chain_settings):
    """headroom.py:292-354 restated with one render per scale, stopping at the first safe one."""
    flat = headroom.flatten_chain_settings(chain_settings)
    flat.pop("return_output_audio")
    sims = []
    for cand in headroom.scaled_candidates(eq_settings):
        sim = _oracle_batch([audio], fs, [(headroom.bands_from_settings(cand), flat)])[0]
        sim["simulation_backend"] = "rust"
        sims.append(sim)
        if headroom.is_headroom_safe(sim):
            break
    return len(sims) - 1, sims


def test_batched_headroom_validation_equals_the_sequential_walk():
    """test_auto_eq.py:968-1001 shape: 0.62 sine at 5 kHz, +9 dB on band 6 -> scaled down, safe."""
    x = (0.62 * np.sin(2 * np.pi * 5000.0 * np.arange(9600) / FS)).astype(np.float32)
    base = {"band_freqs": list(abi.DEFAULT_FREQUENCIES), "band_qs": [1.41] * 10, "band_gains": [0.0] * 10}
    hot = dict(base, band_gains=[0, 0, 0, 0, 0, 0, 9.0, 0, 0, 0])
    flat_ok = dict(base)
    broken = dict(base, band_gains=[1.0, 2.0])
    chain = {"compressor": {"enabled": False}, "limiter": {"enabled": True, "ceiling_db": -0.5}}
    results = headroom.apply_headroom_validation_batch(x, FS, [hot, flat_ok, broken], chain, simulate_batch=_oracle_batch)
    index, sims = _sequential_reference_walk(x, FS, hot, chain)
    r = results[0]
    assert r["headroom_gain_scale"] == headroom.HEADROOM_SCALES[index] < 1.0
    assert r["headroom_safe"] and r["headroom_validation"]["status"] == "safe"
    assert r["headroom_validation"]["after"]["pre_limiter_true_peak_headroom_db"] >= 1.0
    assert r["validation_confidence"] == 0.72
    assert np.allclose(r["band_gains"], np.asarray(hot["band_gains"]) * headroom.HEADROOM_SCALES[index])
    assert results[1]["headroom_gain_scale"] == 1.0 and results[1]["headroom_safe"]
    assert results[2] == broken  # headroom.py:303-304: malformed settings pass through untouched
    # decision margins: the walk examined scales 0 .. index; none of these renders sits within 0.01 dB of a threshold
    margins = r["headroom_validation"]["decision_margins"]
    assert margins["renders_examined"] == index + 1 and margins["near_threshold"] == 0
    worst = min(abs(v) for k in range(index + 1) for v in headroom.headroom_margins_db(sims[k]).values())
    assert margins["smallest_abs_margin_db"] == worst >= headroom.DECISION_MARGIN_DB


def test_decision_margins_flag_renders_next_to_a_threshold():
    safe = {"pre_limiter_true_peak_headroom_db": 1.004, "limiter_gain_reduction_db": 0.2, "true_peak_limiter_gain_reduction_db": 0.1}
    unsafe = {"pre_limiter_true_peak_headroom_db": 3.0, "limiter_gain_reduction_db": 1.5, "true_peak_limiter_gain_reduction_db": 0.1}
    m = headroom.headroom_margins_db(safe)
    assert abs(m["pre_limiter_true_peak_headroom_db"] - 0.004) < 1e-12 and m["limiter_gain_reduction_db"] == 0.8
    report = headroom.decision_margin_report([unsafe, safe, unsafe], examined=1)
    assert report["renders_examined"] == 2 and report["near_threshold"] == 1
    assert report["smallest_at"] == {"scale_index": 1, "metric": "pre_limiter_true_peak_headroom_db"}


def test_abstain_when_no_scale_is_safe():
    x = (0.99 * np.sign(np.sin(2 * np.pi * 900.0 * np.arange(9600) / FS))).astype(np.float32)  # clipped square wave
    base = {"band_freqs": list(abi.DEFAULT_FREQUENCIES), "band_qs": [1.41] * 10, "band_gains": [6.0] * 10}
    r = headroom.apply_headroom_validation_batch(x, FS, [base], {"compressor": {"enabled": False}},
                                                 simulate_batch=_oracle_batch)[0]
    assert r["headroom_gain_scale"] == 0.0 and not r["headroom_safe"]
    assert r["headroom_validation"]["status"] == "risk"
    assert r["validation_confidence"] == 0.42 and r["analysis_confidence"] == 0.58


def test_large_compressor_grids_are_cut_along_passages():
    """build_sweep's piece planner (cut_stream_group): every stream lands in exactly one piece, pieces hold at most
    the limit and whole multiples of 32 streams (but the last), a passage's streams stay together and in order, and
    groups that are small or do not share their prefix stay whole."""
    from tests import hostsim
    n_pass, n_cand = 8, 1000  # pair i = candidate i // n_pass, passage i % n_pass, as in a full cross product
    passage = np.arange(n_pass * n_cand, dtype=np.uint32) % n_pass
    one_eq = np.zeros(passage.size, dtype=np.uint32)
    piece, pos, n = hostsim.cut_group(passage, one_eq, 2048)
    assert n == 4 and piece.max() == 3 and (pos != 0xFFFFFFFF).all()
    sizes = np.bincount(piece)
    assert sizes.sum() == passage.size and sizes.max() <= 2048 and all(k % 32 == 0 for k in sizes[:-1])
    order = np.lexsort((pos, piece))  # streams in execution order
    assert (np.diff(passage[order].astype(np.int64)) >= 0).all()  # passage by passage
    for p in range(n_pass):
        mine = order[passage[order] == p]
        assert (np.diff(mine.astype(np.int64)) > 0).all()  # stable inside a passage
    # at most two pieces see a given passage boundary: a piece spans few passages, so it shares as much as the group
    assert max(len(set(passage[piece == k])) for k in range(n)) <= 3
    # small group / every stream its own EQ / no limit: one piece in the caller's order
    for args in ((passage[:1500], one_eq[:1500], 2048), (passage, np.arange(passage.size, dtype=np.uint32), 2048),
                 (passage, one_eq, 1 << 30)):
        piece, pos, n = hostsim.cut_group(*args)
        assert n == 1 and (piece == 0).all() and (pos == np.arange(args[0].size)).all()
    # a limit that is not a multiple of 32 still bounds the pieces
    piece, pos, n = hostsim.cut_group(passage, one_eq, 1000)
    assert np.bincount(piece).max() <= 1000 and np.bincount(piece).sum() == passage.size


def test_c_abi_partitioner_equals_the_python_one():
    """afsim_multi_partition (what afsim_multi_chain_sweep shards with, one process for all GPUs) and
    sharding.plan_shards (one process per GPU under torchrun) are the same deterministic rule."""
    from audio_forge_b200 import native, workloads
    cands = workloads.full_chain_candidates(96, seed=2)
    lens = [30000, 48000, 12345]
    rng = np.random.default_rng(0)
    pp = rng.integers(0, 3, size=500).astype(np.uint32)
    pc = rng.integers(0, 96, size=500).astype(np.uint32)
    for world in (1, 2, 3, 8):
        owner = native.partition(cands, lens, pp, pc, world)
        shards = sharding.plan_shards(cands, pp, pc, lens, world)
        for r in range(world):
            assert np.array_equal(np.nonzero(owner == r)[0], shards[r])


def test_partition_keeps_a_candidate_s_passages_together_and_still_spreads_few_candidates():
    from audio_forge_b200 import native, workloads
    cands = workloads.full_chain_candidates(256, seed=3)
    n_pass = 8
    pc = np.repeat(np.arange(256), n_pass).astype(np.uint32)
    pp = np.tile(np.arange(n_pass), 256).astype(np.uint32)
    lens = [48000] * n_pass
    shards = sharding.plan_shards(cands, pp, pc, lens, 8)
    assert np.array_equal(np.sort(np.concatenate(shards)), np.arange(256 * n_pass))
    sizes = [s.size for s in shards]
    assert max(sizes) - min(sizes) <= 2 * n_pass
    for s in shards:  # whole candidates: every candidate of a rank comes with all its passages
        held, counts = np.unique(pc[s], return_counts=True)
        assert np.all(counts == n_pass) and held.size * n_pass == s.size
    costs = sharding.stream_costs(cands, pc, np.asarray(lens, dtype=np.float64)[pp])
    loads = np.array([costs[s].sum() for s in shards])
    assert loads.max() / loads.min() < 1.02
    owner = native.partition(cands, lens, pp, pc, 8)
    for r in range(8):
        assert np.array_equal(np.nonzero(owner == r)[0], shards[r])
    # one candidate over many passages (BASELINE config 4): pieces of ceil(n / (16 world)) streams reach every rank
    one = workloads.full_chain_candidates(1, seed=0)
    pp1 = np.arange(1000).astype(np.uint32)
    shards1 = sharding.plan_shards(one, pp1, np.zeros(1000, dtype=np.uint32), [48000] * 1000, 8)
    assert all(120 <= s.size <= 130 for s in shards1)
    owner1 = native.partition(one, [48000] * 1000, pp1, np.zeros(1000, dtype=np.uint32), 8)
    for r in range(8):
        assert np.array_equal(np.nonzero(owner1 == r)[0], shards1[r])


def test_shard_streams_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    costs = rng.choice([50.0, 56.0, 80.0], size=1000) * 480000
    shards = sharding.shard_streams(costs, 8)
    assert np.array_equal(np.sort(np.concatenate(shards)), np.arange(1000))
    loads = np.array([costs[s].sum() for s in shards])
    assert loads.max() / loads.min() < 1.01


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, tmp: str):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        passages, cands, pp, pc = _sweep_problem()
        # the oracle stands in for the rank's GPU render (Simulator.chain_sweep on a B200 rank)
        full = sharding.sharded_chain_sweep(pyoracle.chain_sweep, passages, FS, cands, pp, pc)
        np.save(os.path.join(tmp, f"rank{rank}.npy"), sharding.metrics_to_bytes(full, pp.size))
    finally:
        dist.destroy_process_group()


def _sweep_problem():
    passages = [speech_like(4800, seed=k, level=0.8) for k in range(2)]
    bands = abi.default_bands()
    cand_list = [candidate(bands, compressor_threshold_db=t, compressor_ratio=r) for t in (-30.0, -20.0, -10.0) for r in (2.0, 5.0)]
    cand_list.append(candidate(abi.typed_bands([("high_pass", 90, 0, 0.7, 48, True)] + [("bell", f, 1.0, 1.0, 12, True)
                                                for f in abi.DEFAULT_FREQUENCIES[1:]]), use_typed_bands=True))
    cands = candidate_array(cand_list)
    pp = np.array([p for c in range(len(cand_list)) for p in range(2)], dtype=np.uint32)
    pc = np.array([c for c in range(len(cand_list)) for p in range(2)], dtype=np.uint32)
    return passages, cands, pp, pc


def test_two_rank_gloo_gather_reassembles_the_single_process_result(tmp_path):
    """world_size 2 on CPU: each rank renders its shard, the all-gather puts every struct back in caller order."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    passages, cands, pp, pc = _sweep_problem()
    want = pyoracle.chain_sweep(passages, FS, cands, pp, pc)
    for rank in range(2):
        got = sharding.bytes_to_metrics(np.load(tmp_path / f"rank{rank}.npy"))
        for i in range(pp.size):
            assert metric_mismatches(want[i], got[i]) == {}, (rank, i)
