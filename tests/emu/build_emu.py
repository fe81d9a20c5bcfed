"""TEST INFRASTRUCTURE: compiles the kernel sources against tests/emu/cuda_emu.h with g++ into
tests/emu/_build/libvpho_emu.so so `-m "not gpu"` tests can execute the (non-tensor-core) kernels at toy sizes
on a CPU-only machine.  Never used by the product."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "vpho_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libvpho_emu.so")
# kernels that need real tcgen05/TMA hardware are excluded from the emulated build
EXCLUDE = {"scorenet_tc.cu", "mano_tc.cu", "producers_tc.cu", "runtime.cu"}


def sources():
    return [s for s in sorted(glob.glob(os.path.join(CSRC, "*.cu"))) if os.path.basename(s) not in EXCLUDE]


def build(force: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + \
        [os.path.join(HERE, "cuda_emu.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    flags = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-DVPHO_EMU", "-include",
             os.path.join(HERE, "cuda_emu.h"), "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-w"]
    procs, objs = [], []
    for s in sources():
        o = os.path.join(OUT, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        procs.append((s, subprocess.Popen(["g++"] + flags + ["-x", "c++", "-c", s, "-o", o], stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
    bad = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(f"--- g++ {os.path.basename(s)}\n{out}\n")
            bad = True
    if bad:
        raise RuntimeError("emu build failed")
    r = subprocess.run(["g++", "-shared", "-o", LIB] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("emu link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
