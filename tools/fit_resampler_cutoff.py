"""Solves the f32 `f_cutoff` that rubato 0.14's `calculate_cutoff(sinc_len, window)` returns for the three resampler
configurations the reference ships / evaluates (rust-core/src/audio/processor/resampling.rs:131-156,
python/tools/evaluate_resampler_quality.py), from the reference's OWN published measurements of the real crate
(evaluation/resampler-quality-report.json).  rubato is not vendored in the reference tree and its fitted cutoff
polynomial cannot be restated from memory; but the cutoff is a single f32, and one cutoff-sensitive published
measurement per configuration (a stop-band tone attenuation or the 20 kHz pass-band gain, printed to 1e-14 dB) selects
exactly one f32: the neighbouring f32 values miss the published number by ~1e-5 dB, the solution hits it to ~1e-12.
The remaining ~30 published numbers per configuration are then the CHECK (tools/gen_resampler_golden.py).

Run in the build container; prints the `KNOWN_CUTOFFS` entries of oracle/resampler_oracle.py.
"""
import json
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import resampler_oracle as R  # noqa: E402

REPORT = Path("/root/reference/evaluation/resampler-quality-report.json")


def sine(sr, f, dur):  # evaluate_resampler_quality.py:112-115
    n = int(round(sr * dur))
    return 0.5 * np.sin(2.0 * np.pi * f * (np.arange(n, dtype=np.float64) / sr))


def steady(v, sr):  # :118-122
    m = min(int(round(0.25 * sr)), max(0, v.size // 4))
    return v[m:-m] if m else v


def rms(v):
    return float(np.sqrt(np.mean(np.square(v, dtype=np.float64))))


def db_ratio(a, b):
    return 20.0 * math.log10(a / max(b, 1e-15)) if a > 0 else -300.0


def tone_gain(fc, f, rate_in, rate_out, sinc_len, window, dur):
    src = sine(rate_in, f, dur)
    out, _, expected, _ = R.simulate_product_resampler(src, rate_in, rate_out, 1024, sinc_len, window, f_cutoff=fc)
    return db_ratio(rms(steady(out[:expected], rate_out)), rms(steady(src, rate_in)))


def solve(fn, target, a, b):
    fa, fb = fn(a) - target, fn(b) - target
    assert fa * fb < 0, (a, fa, b, fb)
    c = a
    for _ in range(60):
        c = a - fa * (b - a) / (fb - fa)
        fc = fn(c) - target
        if abs(fc) < 1e-11 or float(np.float32(a)) == float(np.float32(b)):
            break
        if fa * fc < 0:
            b, fb = c, fc
        else:
            a, fa = c, fc
    return c


def neighbours(fn, target, c):
    mid = np.float32(c)
    rows = []
    for k in (-1, 0, 1):
        v = mid
        for _ in range(abs(k)):
            v = np.nextafter(v, np.float32(2.0 if k > 0 else 0.0))
        rows.append((float(v), fn(float(v)) - target))
    return rows


def main():
    rep = json.loads(REPORT.read_text())
    alts = {a["configuration"]["identifier"]: a for a in rep["alternatives"]}
    jobs = [
        # (key, measurement, published value, bracket)
        ((128, "blackman"), lambda fc: tone_gain(fc, 22500.0, 48000, 44100, 128, "blackman", 2.0),
         rep["product"]["measurements"]["downsample_alias"]["worst_alias_db"], (0.9526, 0.9530)),
        ((128, "blackman_harris_squared"), lambda fc: -tone_gain(fc, 20000.0, 44100, 48000, 128, "blackman_harris_squared", 1.5),
         alts["legacy-blackman-harris-squared-128"]["measurements"]["passband_and_offline_reference"][0]["max_absolute_error_db"],
         (0.80, 0.95)),
        ((256, "blackman_harris_squared"), lambda fc: -tone_gain(fc, 20000.0, 48000, 44100, 256, "blackman_harris_squared", 1.5),
         alts["high-rejection-blackman-harris-squared-256"]["measurements"]["passband_and_offline_reference"][1]["max_absolute_error_db"],
         (0.944, 0.95)),
    ]
    only = sys.argv[1:]
    for key, fn, target, bracket in jobs:
        if only and str(key[0]) + key[1] not in only:
            continue
        c = solve(fn, target, *bracket)
        print(key, "root", repr(c))
        for v, r in neighbours(fn, target, c):
            print("    f32", repr(v), "residual dB", r)


if __name__ == "__main__":
    main()
