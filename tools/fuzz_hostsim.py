"""CPU soak at the sample rates where the de-esser's band edges reach Nyquist and intermediate stages can go
non-finite: the product's stage bodies walked on the host (tests/hostsim) against the oracle.
usage: fuzz_hostsim.py [cases [seed]]"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audio_forge_b200 import abi  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests import hostsim  # noqa: E402
from tests.cases import metric_mismatches  # noqa: E402
from tests.signals import speech_like  # noqa: E402
from tools.fuzz_parity import random_case  # noqa: E402


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    failures, nonfinite = [], 0
    for i in range(cases):
        _, _, bands, overrides = random_case(rng)
        fs = float(rng.choice([8000.0, 11025.0, 22050.0, 192000.0]))
        overrides.update(deesser_enabled=True, input_stage="none", compressor_auto_makeup_enabled=False)
        n = int(rng.integers(2000, 20000))
        x = speech_like(n, seed=int(rng.integers(1 << 30)), fs=fs, level=float(rng.uniform(0.05, 1.2)))
        settings = abi.make_settings(**overrides)
        cand = abi.AfCandidate()
        for b in range(abi.NUM_BANDS):
            cand.bands[b] = bands[b]
        cand.settings = settings
        cands = (abi.AfCandidate * 1)(cand)
        try:
            m0, a0, _ = pyoracle.chain_render(x, fs, bands, settings, return_audio=True)
        except pyoracle.OracleError:
            continue
        for split in (0, 1):
            try:
                m1, a1, _ = hostsim.chain_sweep([x], fs, cands, [0], [0], split=split, want_audio=True)
            except hostsim.HostsimError as e:
                failures.append({"case": i, "fs": fs, "split": split, "error": str(e)})
                continue
            bad = metric_mismatches(m0, m1[0])
            same_audio = np.array_equal(a0, a1[0], equal_nan=True)
            if bad or not same_audio:
                failures.append({"case": i, "fs": fs, "split": split, "bad": {k: list(map(float, v)) for k, v in bad.items()},
                                 "audio_equal": bool(same_audio)})
        if not np.all(np.isfinite(a0)):
            nonfinite += 1
    print(json.dumps({"cases": cases, "seed": seed, "failures": len(failures), "oracle_nonfinite_audio": nonfinite, "first": failures[:4]}))


if __name__ == "__main__":
    main()
