"""Turn an `ncu --set full --import-source on` report into the compact text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>.txt

Needs only the ncu CLI (no GPU): raw metrics of every captured launch + warp-stall totals + opcode histogram from the
source page of the first launch.
"""
import csv
import io
import subprocess
import sys
from collections import Counter

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of {rep}")
    for r in rows[2:]:
        print(f"\n## launch id {r[0]}: {r[hdr.index('Kernel Name')]}")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"{m:72s} {r[i]:>22s} {units[i]}")
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        hdr = src[1]
        ix = {h: i for i, h in enumerate(hdr)}

        def f(r, k):
            try:
                return float(r[ix[k]])
            except (ValueError, KeyError, IndexError):
                return 0.0
        data = [r for r in src[2:] if len(r) == len(hdr)]
        tot = sum(f(r, "# Samples") for r in data) or 1.0
        inst = sum(f(r, "Instructions Executed") for r in data) or 1.0
        print(f"\n## warp-stall samples, first launch ({int(tot)} samples, {len(data)} SASS lines)")
        stalls = {h: sum(f(r, h) for r in data) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
        for s, v in sorted(stalls.items(), key=lambda x: -x[1])[:10]:
            print(f"{s:28s} {100 * v / tot:6.1f} %")
        ops, smp = Counter(), Counter()
        for r in data:
            text = r[ix["Source"]].split()
            if not text:
                continue
            op = text[1] if text[0].startswith("@") and len(text) > 1 else text[0]
            ops[op] += f(r, "Instructions Executed")
            smp[op] += f(r, "# Samples")
        print("\n## opcode histogram (share of executed warp instructions / share of stall samples)")
        for op, v in ops.most_common(24):
            print(f"{op:24s} {100 * v / inst:6.1f} %   {100 * smp[op] / tot:6.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
