"""Build libvitlens_b200.so (sm_100a) in-tree with nvcc.  No torch extension machinery: the
library is a plain C-ABI shared object (include/vitlens_b200.h) loaded with ctypes."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
LIB_PATH = os.path.join(HERE, "libvitlens_b200.so")
OBJ_DIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--use_fast_math",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-DVL_EXPORTS",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(os.path.dirname(HERE)), "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def needs_build() -> bool:
    return not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < _deps_mtime()


def build(force: bool = False, verbose: bool = False, ptxas_v: bool = False) -> str:
    """Compile (if stale) under an exclusive file lock: every rank of a torchrun launch may call this at the same moment; one
    builds, the others wait and then find the library up to date."""
    import fcntl

    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another process built it while this one waited
                return LIB_PATH
            return _build_locked(force, verbose, ptxas_v)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool, ptxas_v: bool) -> str:
    nvcc = _nvcc()
    hdr_m = max(os.path.getmtime(os.path.join(r, f))
                for r in (CSRC, os.path.join(os.path.dirname(os.path.dirname(HERE)), "include"))
                for f in os.listdir(r) if f.endswith((".cuh", ".h")))

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_m):
            return obj, ""
        extra = os.environ.get("VL_NVCC_EXTRA", "").split()  # bring-up switches, e.g. -DVL_FWD2_TIMELINE
        cmd = [nvcc, *NVCC_FLAGS, *extra, *(["-Xptxas", "-v"] if ptxas_v else []), "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, sources()))
    objs = [o for o, _ in results]
    if verbose or ptxas_v:
        for _, log in results:
            if log:
                print(log, file=sys.stderr)
    # static cudart: the .so must load (and export its symbols) on a box without a GPU / libcuda
    tmp = LIB_PATH + f".tmp{os.getpid()}"  # link beside the target and rename: a concurrent loader never maps a half-written file
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-cudart", "static", "-Xcompiler", "-fPIC"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, ptxas_v="-v" in sys.argv))
