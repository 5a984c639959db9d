"""In-tree build of libb200ret.so (the C-ABI CUDA library) with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the resulting .so is git-ignored
but travels to the GPU box with the repo snapshot.  No torch headers are involved: the library's ABI is
plain C (include/b200ret.h) and the Python host side binds it with ctypes.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
REPO = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libb200ret.so")
STAMP_PATH = os.path.join(PKG_DIR, ".libb200ret.stamp")

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fopenmp",      # host-side result writer (csrc/run_writer.cpp) formats queries in parallel
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200ret.so cannot be built (there is no CPU fallback)")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(REPO, "include", "b200ret.h"))
    for path in files:
        h.update(path.encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as f:
        return f.read().strip() == _digest()


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libb200ret.so unless an up-to-date build exists. Returns the .so path."""
    if not force and is_current():
        return LIB_PATH
    import fcntl
    with open(os.path.join(PKG_DIR, ".libb200ret.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)          # ranks of one torchrun job must not compile into the same file concurrently
        if not force and is_current():            # another process built it while we waited
            return LIB_PATH
        tmp = f"{LIB_PATH}.tmp{os.getpid()}"
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + _sources()
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        if verbose:
            sys.stderr.write(proc.stderr)
        os.replace(tmp, LIB_PATH)                 # readers never see a torn .so
        with open(STAMP_PATH, "w") as f:
            f.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
