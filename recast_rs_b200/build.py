"""Builds recast_rs_b200/librecast_b200.so (the C-ABI library) with nvcc for sm_100a, in-tree.

Flags follow SURVEY.md Appendix A: no FMA contraction, IEEE div/sqrt, no flush-to-zero, so the
float clipper reproduces the reference's CPU arithmetic bit for bit.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librecast_b200.so")
SOURCES = ["rcb_api.cu", "rcb_multi.cu", "k_scan.cu", "k_raster.cu", "k_raster2.cu", "k_tile.cu", "k_tilecache.cu", "k_area.cu", "k_wire.cu", "k_lz4.cu", "k_regions.cu", "k_filter.cu", "k_compact.cu", "k_dist.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """debug: librecast_b200_dbg.so, the same sources with -DRCB_BOUNDS (checked accessors, csrc/rcb_bounds.cuh)."""
    nvcc = _nvcc()
    hdrs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))  # every header: k_regions_impl.cuh too
    hdrs += [os.path.join(os.path.dirname(HERE), "include", "recast_b200.h"), os.path.abspath(__file__)]
    objdir = os.path.join(CSRC, "build_dbg" if debug else "build")
    out = OUT.replace(".so", "_dbg.so") if debug else OUT
    flags = NVCC_FLAGS + (["-DRCB_BOUNDS"] if debug else [])
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            jobs.append([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=6) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    if jobs or force or _stale(out, objs):
        run([nvcc, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-pthread"])
    return out


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv, debug=True))
