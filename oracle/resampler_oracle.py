"""ORACLE (test infrastructure only -- never imported by the product): CPU restatement of the reference's
`simulate_product_resampler` (rust-core/src/audio/processor/resampling.rs:170-262).

The reference builds `rubato::SincFixedIn::<f64>` (third-party crate **rubato 0.14**, rust-core/Cargo.toml:14; NOT
vendored under /root/reference, no network) with cubic interpolation between 256 oversampled windowed-sinc phases
(resampling.rs:140-156) and streams the signal through it in `chunk_size` blocks, a zero-padded partial block and
zero-input flush blocks (resampling.rs:228-259).  This file restates rubato 0.14's published algorithm:

* `make_window` / `make_sincs` (rubato `windows.rs`, `sinc.rs`): periodic cosine-sum window over sinc_len x 256 points,
  `w[x] * sinc((x - tot/2) * f_cutoff / 256)`, normalised to unit DC gain per phase, stored phase-reversed
  (`sincs[255 - n][p] = y[256 p + n] / sum`);
* `make_interpolator`: the cutoff is scaled by the ratio when downsampling (f32 arithmetic);
* `SincFixedIn::process_into_buffer` (rubato `asynchro.rs`): `last_index = -sinc_len/2`, per block
  `end_idx = chunk - (sinc_len + 1) - ceil(t_ratio)`, `while idx < end_idx { idx += t_ratio; ... }`,
  `last_index = idx - chunk`; `get_nearest_times_4` (sub-phases frac-1 .. frac+2 with carries) and `interp_cubic`;
* `output_delay() = floor(sinc_len / 2 * ratio)`.

`calculate_cutoff` (rubato `sinc.rs`) is a fitted cubic in 1/sinc_len whose coefficients cannot be recalled without the
source; instead the three configurations the reference ships and evaluates (resampling.rs:131-138 and
python/tools/evaluate_resampler_quality.py) carry their f32 cutoffs in `KNOWN_CUTOFFS`, each SOLVED from the reference's
own published measurements (evaluation/resampler-quality-report.json, produced with the real crate) by
`tools/fit_resampler_cutoff.py`: the value is the f32 whose rendered measurements reproduce the published ones.
Other (sinc_len, window) pairs raise.  Parity status: **pinned on published outputs of the real crate** (every
measurement of the report reproduced through the reference's own tool code, tests/test_oracle_resampler_report.py);
the crate's SIMD summation order is unknown, so agreement is to rounding (~1e-13 relative), not bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

OVERSAMPLING = 256
MAX_CHUNK = 1024  # RESAMPLER_CHUNK_SIZE

# f32 cutoffs of `calculate_cutoff(sinc_len, window)`, solved from the published report (tools/fit_resampler_cutoff.py)
KNOWN_CUTOFFS: dict[tuple[int, str], float] = {
    (128, "blackman"): 0.9527542591094971,                 # product default (resampling.rs:131-138)
    (128, "blackman_harris_squared"): 0.8947277069091797,  # the tool's "legacy" alternative
    (256, "blackman_harris_squared"): 0.9470546841621399,  # the tool's "high-rejection" alternative
}

WINDOWS = ("blackman_harris", "blackman_harris_squared", "blackman", "blackman_squared", "hann", "hann_squared")


def make_window(npoints: int, name: str) -> np.ndarray:
    """rubato windows.rs: periodic (denominator npoints) cosine-sum windows; *_squared = elementwise square."""
    x = np.arange(npoints, dtype=np.float64)
    n = float(npoints)
    pi = math.pi
    base = name[:-8] if name.endswith("_squared") else name
    if base == "blackman_harris":
        w = 0.35875 - 0.48829 * np.cos(2.0 * pi * x / n) + 0.14128 * np.cos(4.0 * pi * x / n) - 0.01168 * np.cos(6.0 * pi * x / n)
    elif base == "blackman":
        w = 0.42 - 0.5 * np.cos(2.0 * pi * x / n) + 0.08 * np.cos(4.0 * pi * x / n)
    elif base == "hann":
        w = 0.5 - 0.5 * np.cos(2.0 * pi * x / n)
    else:
        raise ValueError(f"unsupported resampler window {name!r}")
    return w * w if name.endswith("_squared") else w


def make_sincs(sinc_len: int, f_cutoff: float, window: str, factor: int = OVERSAMPLING) -> np.ndarray:
    """rubato sinc.rs make_sincs: [factor][sinc_len] table, phase-reversed, unit DC gain per phase on average."""
    tot = sinc_len * factor
    w = make_window(tot, window)
    arg = (np.arange(tot, dtype=np.float64) - float(tot // 2)) * float(f_cutoff) / float(factor)
    with np.errstate(invalid="ignore", divide="ignore"):
        s = np.where(arg == 0.0, 1.0, np.sin(arg * math.pi) / (arg * math.pi))
    y = w * s
    total = 0.0
    for v in y:  # sequential f64 sum as the crate's loop
        total += float(v)
    total /= float(factor)
    table = np.empty((factor, sinc_len), dtype=np.float64)
    yy = (y / total).reshape(sinc_len, factor)  # yy[p][n] = y[factor p + n] / sum
    table[:, :] = yy.T[::-1, :]                 # sincs[factor - n - 1][p]
    return table


def effective_cutoff(f_cutoff: float, ratio: float) -> float:
    """rubato make_interpolator: `if ratio >= 1 { f } else { f * ratio as f32 }` in f32."""
    f = np.float32(f_cutoff)
    if ratio >= 1.0:
        return float(f)
    return float(np.float32(f * np.float32(ratio)))


def output_positions(n_in: int, input_rate: int, output_rate: int, chunk_size: int, sinc_len: int, need: int):
    """The interpolation positions `idx` of every produced frame, block by block as SincFixedIn walks them
    (sequential f64 additions, re-based by -chunk per block), until at least `need` frames exist AND the input
    (full blocks, one zero-padded partial block) is consumed.  Returns (block_of_frame, idx) arrays; frame k of
    block b reads chunk coordinate floor(idx) relative to the start of block b."""
    ratio = output_rate / input_rate
    t_ratio = 1.0 / ratio
    end_idx = float(chunk_size - (sinc_len + 1) - math.ceil(t_ratio))
    last = -float(sinc_len // 2)
    blocks, idxs = [], []
    produced = 0
    n_blocks_in = -(-n_in // chunk_size)
    b = 0
    while b < n_blocks_in or produced < need:
        if last < end_idx:
            est = int((end_idx - last) / t_ratio) + 4
            walk = np.cumsum(np.concatenate(([last], np.full(est, t_ratio))))  # ((last + t) + t) + ... in order
            steps, before = walk[1:], walk[:-1]
            k = int(np.count_nonzero(before < end_idx))               # loop test precedes the increment
            assert k < est
            cur = steps[:k]
            if k:
                idxs.append(cur)
                blocks.append(np.full(k, b, dtype=np.int64))
                produced += k
                last = float(cur[-1])
        elif b >= n_blocks_in:
            raise RuntimeError("resampler flush produced no frames before reaching the expected length")
        last = last - float(chunk_size)
        b += 1
    if not idxs:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    return np.concatenate(blocks), np.concatenate(idxs)


def frame_list(n_in: int, input_rate: int, output_rate: int, chunk_size: int, sinc_len: int):
    """Per produced frame: the absolute input sample under tap 0 of phase `sub` (buffer position floor(idx) +
    2 sinc_len of block b holds sample b * chunk + floor(idx)), the phase (get_nearest_times_4's middle-left point) and
    the cubic's abscissa `frac - floor(frac)`, frac = idx * 256."""
    ratio = output_rate / input_rate
    delay = int(sinc_len / 2 * ratio)
    expected = int(math.floor(n_in * float(output_rate) / float(input_rate) + 0.5))
    blk, idx = output_positions(n_in, input_rate, output_rate, chunk_size, sinc_len, expected + delay)
    fl = np.floor(idx)
    base = blk * chunk_size + fl.astype(np.int64)
    frac_f = idx * float(OVERSAMPLING)
    frac_off = frac_f - np.floor(frac_f)
    sub = np.floor((idx - fl) * float(OVERSAMPLING)).astype(np.int64)
    return base, sub, frac_off


def simulate_product_resampler(samples, input_rate, output_rate, chunk_size=1024, sinc_len=None, window=None,
                               f_cutoff=None, block_outputs=8192):
    """(output, delay, expected_frames, block_times_ns) as resampling.rs:179-262; block_times_ns is empty (a wall-clock
    measurement, not part of the rendered result)."""
    x = np.asarray(samples, dtype=np.float64)
    if input_rate <= 0 or output_rate <= 0:
        raise ValueError("sample rates must be positive")
    if not 1 <= chunk_size <= MAX_CHUNK:
        raise ValueError(f"chunk_size must be between 1 and {MAX_CHUNK}")
    sinc_len = 128 if sinc_len is None else int(sinc_len)
    if not 32 <= sinc_len <= 2048 or sinc_len & (sinc_len - 1):
        raise ValueError("sinc_len must be a power of two between 32 and 2048")
    window = "blackman" if window is None else window
    if window not in WINDOWS:
        raise ValueError(f"unsupported resampler window {window!r}")
    if not np.isfinite(x).all():
        raise ValueError("samples must be finite")
    if f_cutoff is None:
        if (sinc_len, window) not in KNOWN_CUTOFFS:
            raise NotImplementedError(f"no pinned calculate_cutoff value for sinc_len {sinc_len}, window {window}")
        f_cutoff = KNOWN_CUTOFFS[(sinc_len, window)]
    ratio = output_rate / input_rate
    table = make_sincs(sinc_len, effective_cutoff(f_cutoff, ratio), window)
    delay = int(sinc_len / 2 * ratio)  # output_delay(): (sinc_len / 2) as f64 * ratio, truncated
    expected = int(math.floor(x.size * float(output_rate) / float(input_rate) + 0.5))
    base, sub, frac_off = frame_list(x.size, input_rate, output_rate, chunk_size, sinc_len)
    n_out = base.size
    pad_lo = 2 * sinc_len + 2
    total_in = (int(base[-1]) if n_out else 0) + chunk_size + 2 * sinc_len
    xp = np.zeros(pad_lo + total_in + sinc_len + 2, dtype=np.float64)
    xp[pad_lo:pad_lo + x.size] = x
    out = np.zeros(n_out, dtype=np.float64)
    if not x.any():  # silence in, silence out (the tool's 60 s sample-count case): skip the dot products
        return out, delay, expected, []
    taps = np.arange(sinc_len, dtype=np.int64)
    for s0 in range(0, n_out, block_outputs):
        s1 = min(n_out, s0 + block_outputs)
        pts = []
        for d in (-1, 0, 1, 2):
            sd = sub[s0:s1] + d
            carry = np.floor_divide(sd, OVERSAMPLING)
            sd = sd - carry * OVERSAMPLING
            pos = base[s0:s1] + carry + pad_lo
            seg = xp[pos[:, None] + taps[None, :]]
            pts.append(np.einsum("ij,ij->i", seg, table[sd]))
        y0, y1, y2, y3 = pts
        a0 = y1
        a1 = -(1.0 / 3.0) * y0 - 0.5 * y1 + y2 - (1.0 / 6.0) * y3
        a2 = 0.5 * (y0 + y2) - y1
        a3 = 0.5 * (y1 - y2) + (1.0 / 6.0) * (y3 - y0)
        xo = frac_off[s0:s1]
        x2 = xo * xo
        out[s0:s1] = a0 + a1 * xo + a2 * x2 + a3 * (x2 * xo)
    return out, delay, expected, []


def product_resampler_configuration():
    """resampling.rs:263-272."""
    return (128, "blackman", "cubic", OVERSAMPLING, MAX_CHUNK)
