"""dev helper: build a tuning variant of libb200ret.so into variants/<tag>.so with extra -D flags.

    python tools/build_variant.py desc0 -DB200RET_DESC_MODE=0

bench.py / the tests pick a variant up through B200RET_LIB=variants/<tag>.so (see tools/dev_variants.sh)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scaling_retriever_b200 import build  # noqa: E402

tag, flags = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
out = os.path.join(ROOT, "variants", tag + ".so")
cmd = [build._nvcc()] + build.NVCC_FLAGS + flags + ["-o", out] + build._sources()
proc = subprocess.run(cmd, capture_output=True, text=True)
if proc.returncode != 0:
    sys.exit(proc.stdout + proc.stderr)
print(out)
