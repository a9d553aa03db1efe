"""Brief per-launch view of an .ncu-rep: duration, warp instructions, issue rate, pipes, DRAM, top stall reasons."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, k):
        try:
            return float(r[col[k]].replace(",", ""))
        except Exception:
            return float("nan")

    for r in data:
        name = r[col["Kernel Name"]].split("(")[0]
        stalls = sorted(((get(r, h), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in hdr
                         if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h), reverse=True)
        tot = sum(v for v, _ in stalls if v == v) or 1.0
        top = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in stalls[:5])
        print(f"{name}: {get(r, 'gpu__time_duration.sum'):.3f} {r and rows[1][col['gpu__time_duration.sum']]}, "
              f"grid {r[col['launch__grid_size']]} x {r[col['launch__block_size']]}, regs {r[col['launch__registers_per_thread']]}, "
              f"warp instr {get(r, 'smsp__inst_executed.sum'):.3g}, issue active {get(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f}%, "
              f"fp64 pipe {get(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f}%, "
              f"dram rd+wr {get(r, 'dram__bytes_read.sum') + get(r, 'dram__bytes_write.sum'):.1f} {rows[1][col['dram__bytes_read.sum']]}, "
              f"local ld/st {get(r, 'smsp__inst_executed_op_local_ld.sum') + get(r, 'smsp__inst_executed_op_local_st.sum'):.3g}; stalls: {top}")


if __name__ == "__main__":
    main()
