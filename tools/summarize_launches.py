"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel -> markdown table."""
import collections
import csv
import sys


def main(path: str, title: str, command: str) -> None:
    rows = list(csv.DictReader(line for line in open(path) if line.startswith('"')))
    agg: dict = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    total = sum(v[1] for v in agg.values()) or 1.0
    print(f"# {title}\n")
    print(f"Command: `{command}`")
    print("(cold-cache, serialised per-launch times under ncu: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for name, (n, us) in agg.items():
        print(f"| {name} | {n} | {us:.1f} | {us / n:.1f} | {us / total:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
