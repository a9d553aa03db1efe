"""ncu `--page raw --csv` export -> the per-stage table bench.py loads (profiles/ncu_tables.json).

usage: ncu_to_tables.py <raw.csv> <key> [--md out.md]     key = "<workload>:<streams of the first batch>:<chunk>"
Per stage kernel (mean over its captured launches): duration, warp instructions, DRAM bytes read + written, issue-active
and FP64 / FMA pipe utilisation.  Merged into profiles/ncu_tables.json under `key`."""
import csv
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
STAGE_OF = [
    (r"k_input_fanout", "input_fanout"), (r"k_input_cleanup|k_input\b|k_input\(", "input"), (r"k_eq<", "eq"),
    (r"k_comp_r1", "comp_r1"), (r"k_comp_m2", "comp_m2"), (r"k_comp_r3", "comp_r3"), (r"k_comp_m4", "comp_m4"),
    (r"k_comp_r5", "comp_r5"), (r"k_comp_m6", "comp_m6"), (r"k_comp_r7", "comp_r7"), (r"k_lim_m", "lim_m"), (r"k_lim_r", "lim_r"),
    (r"k_tp_fir_in", "tp_fir_in"), (r"k_tp_r", "tp_r"), (r"k_tp_fir_out", "tp_fir_out"), (r"k_tail", "tail"),
    (r"k_de_ra", "de_ra"), (r"k_de_mb", "de_mb"), (r"k_de_rc1a", "de_rc1a"), (r"k_de_mc1b", "de_mc1b"), (r"k_de_rc1c", "de_rc1c"),
    (r"k_de_mc2", "de_mc2"), (r"k_de_rc3", "de_rc3"), (r"k_de_rc", "de_rc"), (r"k_compressor", "compressor"),
    (r"k_limiter", "limiter"), (r"k_output", "output"), (r"k_finalize", "finalize"), (r"k_expand_deesser", "expand_deesser"),
]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0,
        "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def stage_of(kernel: str):
    for pat, name in STAGE_OF:
        if re.search(pat, kernel):
            return name
    return None


def main():
    path, key = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(path, newline="")))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[start], rows[start + 1], rows[start + 2:]

    def col(name):
        return hdr.index(name) if name in hdr else None

    cols = {"us": col("gpu__time_duration.sum"), "instr": col("smsp__inst_executed.sum"), "rd": col("dram__bytes_read.sum"),
            "wr": col("dram__bytes_write.sum"), "issue": col("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fp64": col("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "fma": col("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "l2hit": col("lts__t_sector_hit_rate.pct"), "grid": col("launch__grid_size"), "block": col("launch__block_size"),
            "regs": col("launch__registers_per_thread")}
    kname = col("Kernel Name")

    def val(r, c):
        if c is None or c >= len(r) or r[c] in ("", "n/a"):
            return None
        try:
            return float(r[c].replace(",", "")) * UNIT.get(units[c], 1.0)
        except ValueError:
            return None

    acc = {}
    for r in data:
        if len(r) <= kname:
            continue
        st = stage_of(r[kname])
        if st is None:
            continue
        a = acc.setdefault(st, {"n": 0, "kernel": r[kname].split("(")[0], "grid": r[cols["grid"]] if cols["grid"] is not None else "",
                                "block": r[cols["block"]] if cols["block"] is not None else "",
                                "regs": r[cols["regs"]] if cols["regs"] is not None else "", "sums": {}})
        a["n"] += 1
        for k in ("us", "instr", "rd", "wr", "issue", "fp64", "fma", "l2hit"):
            v = val(r, cols[k])
            if v is not None:
                s = a["sums"].setdefault(k, [0.0, 0])
                s[0] += v
                s[1] += 1
    table = {}
    for st, a in acc.items():
        m = {k: s[0] / s[1] for k, s in a["sums"].items() if s[1]}
        table[st] = {"kernel": a["kernel"], "launches": a["n"], "grid": a["grid"], "block": a["block"], "regs": a["regs"],
                     "us": m.get("us"), "warp_instr": m.get("instr"), "dram_bytes": (m.get("rd", 0.0) + m.get("wr", 0.0)),
                     "issue_active_pct": m.get("issue"), "fp64_pipe_pct": m.get("fp64"), "fma_pipe_pct": m.get("fma"),
                     "l2_hit_pct": m.get("l2hit")}
    out = ROOT / "profiles" / "ncu_tables.json"
    tables = json.loads(out.read_text()) if out.exists() else {}
    tables[key] = table
    out.write_text(json.dumps(tables, indent=1, sort_keys=True) + "\n")
    lines = [f"ncu --set full --clock-control none, {key} (mean per launch)", "",
             "| stage | kernel | grid x block | regs | launches | us | warp instr | DRAM rd+wr MB | issue active % | fp64 pipe % | fma pipe % | L2 hit % |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]

    def f(v, fmt="{:.1f}"):
        return fmt.format(v) if v is not None else "-"
    for st, t in sorted(table.items(), key=lambda kv: -(kv[1]["us"] or 0)):
        lines.append(f"| {st} | {t['kernel']} | {t['grid']} x {t['block']} | {t['regs']} | {t['launches']} | {f(t['us'])} | "
                     f"{f(t['warp_instr'], '{:.3g}')} | {f(t['dram_bytes'] / 1e6 if t['dram_bytes'] is not None else None)} | "
                     f"{f(t['issue_active_pct'])} | {f(t['fp64_pipe_pct'])} | {f(t['fma_pipe_pct'])} | {f(t['l2_hit_pct'])} |")
    text = "\n".join(lines) + "\n"
    if "--md" in sys.argv:
        Path(sys.argv[sys.argv.index("--md") + 1]).write_text(text)
    print(text)


if __name__ == "__main__":
    main()
