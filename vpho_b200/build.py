"""Builds the sm_100a shared library `vpho_b200/csrc/libvpho_b200.so` in-tree with nvcc.

There is exactly one product binary and it is CUDA: no CPU build of the product exists.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vpho_b200", "csrc")
LIB = os.path.join(CSRC, "libvpho_b200.so")
LIB_BOUNDS = os.path.join(CSRC, "libvpho_b200_bounds.so")      # -DVPHO_DEBUG_BOUNDS twin, used by tests/test_bounds_build.py only
LIB_TIMELINE = os.path.join(CSRC, "libvpho_b200_timeline.so")  # -DVPHO_TC_TIMELINE twin, used by tools/diag_*_timeline.py only
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale(lib: str = LIB) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, bounds: bool = False, timeline: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into one shared library (object per file, parallel).
    bounds=True: the index-asserting twin (VPHO_BOUNDS in vpho_common.cuh) -> libvpho_b200_bounds.so."""
    LIB = LIB_BOUNDS if bounds else LIB_TIMELINE if timeline else globals()["LIB"]
    if not force and not _stale(LIB):
        return LIB
    objdir = os.path.join(CSRC, "build_bounds" if bounds else "build_timeline" if timeline else "build")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
             "-I", os.path.join(ROOT, "include"), "-I", CSRC] + ARCH
    if bounds:
        flags += ["-DVPHO_DEBUG_BOUNDS"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    if timeline or os.environ.get("VPHO_TC_TIMELINE"):
        flags += ["-DVPHO_TC_TIMELINE"]
    procs = []
    objs = []
    for s in sources():
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        procs.append((s, subprocess.Popen([NVCC, "-c", s, "-o", o] + flags, stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {os.path.basename(s)} (rc={p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ARCH + ["-lcudart", "-lcuda"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, bounds="--bounds" in sys.argv,
                timeline="--timeline" in sys.argv))
