"""Builds surf_b200/csrc/libsurf_b200.so (sm_100a only) with plain nvcc, in-tree.

    python -m surf_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(CSRC, "libsurf_b200.so")
SOURCES = ["scene.cu", "sample.cu", "sdf_mlp.cu", "blend.cu", "render.cu", "tc_selftest.cu", "sdf_tc2.cu", "blend_tc.cu", "marching.cu", "mesh_clean.cu", "extras.cu", "sdf_smooth.cu", "sdf_smooth_tc.cu", "matching.cu", "volume.cu", "host_rng.cu", "fpn.cu"]
EXTRA = os.environ.get("SURF_NVCC_EXTRA", "").split()
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = nvcc_path()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "surf_b200.h"))
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo", *ARCH, *EXTRA, "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC,
                   "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            print("nvcc failed for %s:\n%s" % (src, out), file=sys.stderr)
        elif verbose and out.strip():
            print(out)
    if failed:
        raise RuntimeError("surf_b200: CUDA build failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", *ARCH, "-o", LIB, *objs, "-lcudart"]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("surf_b200: link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
