"""Builds libsie_b200.so (sm_100a) in-tree with nvcc.  Used by `__graft_entry__.build()`.

The shared object sits next to this file so it travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsie_b200.so")
SOURCES = ["abi.cu", "detrend.cu", "corr.cu", "area.cu", "links.cu", "gp.cu", "ingest.cu"]
# SIE_AREA_TIMERS=1: per-phase clock64 counters of k_area_level (work[4..]); they cost ~6 % of the kernel, so the
# product build leaves them out (tools/prof_sweep.py and tools/prof_one.py need a build with them)
NVCC_FLAGS = (["-DSIE_AREA_PHASE_TIMERS"] if os.environ.get("SIE_AREA_TIMERS") else []) + \
    (["-DSIE_AREA_LDG"] if os.environ.get("SIE_AREA_LDG") else []) + \
    (["-D" + f for f in os.environ.get("SIE_DEFINES", "").split() if f]) + \
    (["-DSIE_PW_SMALL_NQ=" + os.environ["SIE_PW_SMALL_NQ"]] if os.environ.get("SIE_PW_SMALL_NQ") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "sie_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
