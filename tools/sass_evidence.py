"""Per-kernel SASS mnemonic counts of the shipped libb200ret.so (cuobjdump; runs without a GPU):

    python tools/sass_evidence.py > profiles/r02_sass_counts.txt

Evidence that the dense kernel is hand-issued tcgen05 / TMEM / TMA code (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load,
LDTM = tcgen05.ld, UTCBAR = tcgen05.commit) and that no kernel uses library or legacy tensor paths (HMMA/IMMA = mma.sync)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "scaling_retriever_b200", "libb200ret.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTCATOM", "SYNCS", "HMMA", "IMMA", "LDGSTS", "LDG", "LDS", "STS", "ATOMS", "ATOMG",
         "REDG", "RED", "BAR", "WARPSYNC", "SHFL", "VOTE", "MATCH", "FADD", "FMUL", "FFMA"]
fn = None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[fn][w] += 1
print("# SASS mnemonic counts per kernel of scaling_retriever_b200/libb200ret.so (cuobjdump -sass, sm_100a)")
for fn, c in counts.items():
    items = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{fn}: instructions={c['_total']} {items}")
dense = next((c for f, c in counts.items() if "dense_search_kernel" in f), None)
assert dense and dense["UTCHMMA"] and dense["UTMALDG"] and dense["LDTM"], "dense kernel lost its tcgen05/TMA instructions"
assert not any(c["HMMA"] or c["IMMA"] for c in counts.values()), "legacy mma.sync found"
