"""profiles/rNN_launches.csv (ncu launch list with dram__bytes_read/write per launch) -> profiles/rNN_traffic.json:
average DRAM bytes per launch of the two dominant kernels, with the launch counts they were averaged over, for bench.py's
`roofline.traffic` (used only while the kernel shape recorded here still matches the running library).

    python tools/traffic_from_launches.py profiles/r02_launches.csv profiles/r02_traffic.json <bench steps in the capture>
"""
import csv
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(src, dst, steps):
    rows = [r for r in csv.reader(open(src)) if len(r) > 8]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = defaultdict(lambda: defaultdict(float))
    for r in rows[1:]:
        try:
            name = r[ki].split("(")[0].replace("void ", "").split("<")[0]        # template instances count under the kernel's name
            per[name][(int(r[ii]), r[mi])] = float(r[vi].replace(",", ""))
        except ValueError:
            continue
    from scaling_retriever_b200 import _lib
    out = {"source": os.path.relpath(src, ROOT), "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
           "--clock-control none over `bench.py --steps 2 --warmup 3` (tools/profile_r02_launches.sh)",
           "sparse_block_docs": int(_lib.load().b200ret_sparse_block_docs()), "bench_steps_in_capture": steps, "kernels": {}}
    for name in ("sparse_score_kernel", "dense_search_kernel"):
        m = per[name]
        ids = sorted({i for i, _ in m})
        total = sum(m[(i, "dram__bytes_read.sum")] + m[(i, "dram__bytes_write.sum")] for i in ids)
        out["kernels"][name] = {"launches": len(ids), "launches_per_step": len(ids) / steps, "dram_bytes_per_launch": total / max(len(ids), 1),
                                "dram_bytes_per_step": total / steps}
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]))
