"""Randomised parity soak: random sample rates, lengths, bands and settings through libafsim.so against the CPU
oracle (north_star tolerances).  usage: fuzz_parity.py [cases [seed]]   -- prints one JSON summary line."""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_forge_b200 import abi, native  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests.cases import audio_within_tolerance, metric_mismatches  # noqa: E402
from tests.signals import speech_like  # noqa: E402


def random_case(rng):
    fs = float(rng.choice([8000.0, 11025.0, 16000.0, 22050.0, 32000.0, 44100.0, 48000.0, 96000.0]))
    n = int(rng.choice([rng.integers(1, 200), rng.integers(200, 4000), rng.integers(4000, 40000)]))
    nyq = fs / 2.0 - 1.0
    typed = bool(rng.random() < 0.5)
    if typed:
        kinds = ["low_shelf", "bell", "high_shelf", "notch", "high_pass", "low_pass"]
        bands = abi.typed_bands([(str(rng.choice(kinds)), float(rng.uniform(20.0, min(nyq, 18000.0))), float(rng.uniform(-12, 12)),
                                  float(rng.uniform(0.1, 10.0)), int(rng.choice([12, 24, 36, 48])), bool(rng.random() < 0.85))
                                 for _ in range(10)])
    else:
        bands = abi.legacy_bands([(float(rng.uniform(20.0, min(nyq, 18000.0))), float(rng.uniform(-12, 12)), float(rng.uniform(0.1, 10.0)))
                                  for _ in range(10)])
    stage = str(rng.choice(["none", "none", "dc_hp80", "gentle", "strong"])) if fs == 48000.0 else str(rng.choice(["none", "dc_hp80"]))
    auto_makeup = bool(rng.random() < 0.25) and stage in ("none", "dc_hp80")
    overrides = dict(
        use_typed_bands=typed, input_stage=stage, eq_before_deesser=bool(rng.random() < 0.5),
        deesser_enabled=bool(rng.random() < 0.5), deesser_auto_enabled=bool(rng.random() < 0.6),
        deesser_auto_amount=float(rng.uniform(0, 1)), deesser_low_cut_hz=float(rng.uniform(3000, 6000)),
        deesser_high_cut_hz=float(rng.uniform(min(7000.0, 0.8 * nyq), min(12000.0, 1.1 * nyq))), deesser_threshold_db=float(rng.uniform(-50, -10)),
        deesser_ratio=float(rng.uniform(1.5, 10)), deesser_attack_ms=float(rng.uniform(0.5, 10)),
        deesser_release_ms=float(rng.uniform(20, 200)), deesser_max_reduction_db=float(rng.uniform(2, 18)),
        compressor_enabled=bool(rng.random() < 0.8), compressor_threshold_db=float(rng.uniform(-50, -6)),
        compressor_ratio=float(rng.uniform(1.0, 8.0)), compressor_attack_ms=float(rng.uniform(1, 30)),
        compressor_release_ms=float(rng.uniform(40, 400)), compressor_makeup_gain_db=float(rng.uniform(0, 12)),
        compressor_base_release_ms=float(rng.uniform(30, 120)), compressor_adaptive_release=bool(rng.random() < 0.5),
        compressor_sidechain_highpass_enabled=bool(rng.random() < 0.6), compressor_auto_makeup_enabled=auto_makeup,
        compressor_target_lufs=float(rng.uniform(-26, -10)),
        limiter_enabled=bool(rng.random() < 0.8), limiter_ceiling_db=float(rng.uniform(-9, 0)),
        limiter_release_ms=float(rng.uniform(5, 300)), limiter_lookahead_ms=float(rng.choice([0.1, 1.0, 2.0, 5.0, 10.0])),
        limiter_careful_output_enabled=bool(rng.random() < 0.5))
    x = speech_like(max(n, 64), seed=int(rng.integers(1 << 30)), fs=fs, level=float(rng.uniform(0.05, 1.2)))[:n].copy()
    return fs, x, bands, overrides


def sweep_soak(rounds: int, seed: int):
    """Random candidate x passage sweeps (mixed stage sets, lengths, shared passages) against the threaded oracle."""
    from tests.cases import candidate_array
    rng = np.random.default_rng(seed)
    sim = native.Simulator(0)
    failures, pairs_checked = [], 0
    for r in range(rounds):
        fs = 48000.0
        lengths = [int(rng.integers(2000, 30000)) for _ in range(int(rng.integers(1, 4)))]
        passages = [speech_like(n, seed=int(rng.integers(1 << 30)), level=float(rng.uniform(0.2, 1.1))) for n in lengths]
        cand_list = []
        for _ in range(int(rng.integers(8, 90))):
            while True:
                cfs, _, bands, overrides = random_case(rng)
                if cfs == fs:
                    break
            c = abi.AfCandidate()
            for b in range(abi.NUM_BANDS):
                c.bands[b] = bands[b]
            c.settings = abi.make_settings(**overrides)
            cand_list.append(c)
        cands = candidate_array(cand_list)
        n_pass, n_cand = len(passages), len(cand_list)
        pp = np.array([p for c in range(n_cand) for p in range(n_pass)], dtype=np.uint32)
        pc = np.array([c for c in range(n_cand) for p in range(n_pass)], dtype=np.uint32)
        got, _ = sim.chain_sweep(passages, fs, cands)
        want = pyoracle.chain_sweep(passages, fs, cands, pp, pc, n_threads=16)
        for i in range(pp.size):
            bad = metric_mismatches(want[i], got[i], tol_db=0.01)
            pairs_checked += 1
            if bad:
                failures.append({"round": r, "pair": i, "bad": {k: list(v) for k, v in bad.items()}})
    print(json.dumps({"sweep_rounds": rounds, "seed": seed, "pairs": pairs_checked, "failures": len(failures), "first": failures[:3]}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        sweep_soak(int(sys.argv[2]) if len(sys.argv) > 2 else 10, int(sys.argv[3]) if len(sys.argv) > 3 else 1)
        return
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    sim = native.Simulator(0)
    failures = []
    for i in range(cases):
        fs, x, bands, overrides = random_case(rng)
        settings = abi.make_settings(**overrides)
        try:
            m0, a0, _ = pyoracle.chain_render(x, fs, bands, settings, return_audio=True)
        except pyoracle.OracleError as e:
            m0 = None
            err0 = str(e)
        try:
            m1, a1 = sim.chain_render(x, fs, bands, settings, return_audio=True)
        except (ValueError, native.AfsimError) as e:
            if m0 is None:
                continue  # both reject
            if isinstance(e, native.AfsimError) and e.status == abi.AFSIM_UNSUPPORTED and "Nyquist" in e.message:
                continue  # de-esser band edges at / beyond Nyquist: unstable in the reference itself, rejected loudly here
            failures.append({"case": i, "error": str(e), "overrides": overrides, "fs": fs, "n": int(x.size)})
            continue
        if m0 is None:
            failures.append({"case": i, "error": "oracle rejected: " + err0, "fs": fs, "n": int(x.size)})
            continue
        excess = audio_within_tolerance(a0, a1)
        bad = metric_mismatches(m0, m1, tol_db=0.01)
        if excess > 0.0 or bad:
            failures.append({"case": i, "fs": fs, "n": int(x.size), "excess": excess, "bad": {k: list(v) for k, v in bad.items()},
                             "overrides": overrides})
    print(json.dumps({"cases": cases, "seed": seed, "failures": len(failures), "first": failures[:3]}))


if __name__ == "__main__":
    main()
