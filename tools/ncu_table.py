"""Per-kernel table from an ncu --set full report: duration, instruction mix, pipe utilisation, DRAM traffic."""
import csv
import subprocess
import sys


def main(path: str) -> None:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]

    def col(name):
        return hdr.index(name) if name in hdr else None

    cols = {
        "kernel": col("Kernel Name"), "us": col("gpu__time_duration.sum"), "regs": col("launch__registers_per_thread"),
        "grid": col("launch__grid_size"), "block": col("launch__block_size"),
        "inst": col("smsp__inst_executed.sum"), "fp64": col("smsp__inst_executed_pipe_fp64.sum"),
        "fma": col("smsp__inst_executed_pipe_fma.sum"),
        "issue_pct": col("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fp64_pct": col("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "dram_rd": col("dram__bytes_read.sum"), "dram_wr": col("dram__bytes_write.sum"),
        "l2_hit": col("lts__t_sector_hit_rate.pct"), "warps_pct": col("sm__warps_active.avg.pct_of_peak_sustained_active"),
    }
    units = rows[1]
    seen = set()
    print("| kernel | grid x block | regs | us | warp instr | fp64 pipe instr | issue active % | fp64 pipe % | DRAM rd+wr MB | L2 hit % |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for r in rows[2:]:
        name = r[cols["kernel"]].split("(")[0]
        if name in seen:
            continue
        seen.add(name)

        def val(key, scale=1.0):
            i = cols[key]
            if i is None or r[i] == "":
                return float("nan")
            return float(r[i].replace(",", "")) * scale

        def mb(key):
            i = cols[key]
            if i is None:
                return float("nan")
            v = float(r[i].replace(",", ""))
            u = units[i]
            return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)

        us = val("us") * ({"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[cols["us"]], 1.0))
        print(f"| {name} | {r[cols['grid']]} x {r[cols['block']]} | {r[cols['regs']]} | {us:.1f} | {val('inst'):.3g} | "
              f"{val('fp64'):.3g} | {val('issue_pct'):.1f} | {val('fp64_pct'):.1f} | {mb('dram_rd') + mb('dram_wr'):.1f} | {val('l2_hit'):.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
