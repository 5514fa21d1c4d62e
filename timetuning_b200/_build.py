"""Build libtimet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The built library lives next to the sources (timetuning_b200/_lib/) so it travels to the GPU
box with the repo snapshot; it is git-ignored (*.so).  No JIT at import: `build()` is called by
__graft_entry__.build(), by `python -m timetuning_b200._build`, and lazily by _cabi if the .so is
missing or older than its sources.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIBDIR, "libtimet_b200.so")
STAMP = os.path.join(LIBDIR, "build.stamp")

SOURCES = ["common.cu", "sinkhorn.cu", "comm.cu", "misc.cu", "ff_prepare.cu", "ff_select_exact.cu",
           "ff_gather.cu", "ff_tc.cu", "ff_tc3.cu", "scores.cu", "ff_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr", "-diag-suppress", "128"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/timet_b200.h"]
    for f in files:
        path = os.path.join(CSRC, f)
        if os.path.isfile(path):
            h.update(f.encode())
            h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return os.path.isfile(LIB) and os.path.isfile(STAMP) and open(STAMP).read().strip() == source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile and link under an exclusive file lock: under torchrun every rank may get here at once; the first
    one builds, the others wait and then find the library current.  The .so is linked to a temporary name and
    renamed into place, so a concurrent loader never sees a half-written file."""
    if not force and is_current():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIBDIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():          # another process built it while we waited
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed.append(f"--- {src}\n{out}")
        elif verbose or out.strip():
            print(f"--- {src}\n{out}", file=sys.stderr)
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(failed))
    tmp = LIB + f".tmp{os.getpid()}"
    link = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    os.replace(tmp, LIB)
    with open(STAMP + ".tmp", "w") as f:
        f.write(source_hash())
    os.replace(STAMP + ".tmp", STAMP)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
