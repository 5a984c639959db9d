"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list per kernel:
    python profiles/summarize_launches.py gpurun_out/launches.csv [first_launch_id_of_the_timed_step]
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import csv
import sys
from collections import defaultdict


def main(path, first=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 8]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = defaultdict(dict)
    names = {}
    for r in rows[1:]:
        try:
            launches[int(r[ii])][r[mi]] = float(r[vi].replace(",", ""))
            names[int(r[ii])] = r[ki].split("(")[0]
        except ValueError:
            continue
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for i, m in launches.items():
        if i < first:
            continue
        a = agg[names[i]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in agg.values()) or 1.0
    print(f"# {path}: launches with id >= {first}")
    print(f"{'kernel':58s} {'launches':>8s} {'time':>12s} {'share':>7s} {'dram bytes/launch':>18s}")
    for n, (c, t, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{n[:58]:58s} {c:8d} {t:12.0f} {100 * t / total:6.1f}% {b / c:18.3e}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
