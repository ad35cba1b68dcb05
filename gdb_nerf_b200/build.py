"""In-tree build of the C-ABI library (nvcc cross-compiles sm_100a without a GPU).

    python -m gdb_nerf_b200.build [--force] [--verbose]

Output: gdb_nerf_b200/libgdbnerf_b200.so (git-ignored, travels to the GPU box).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from typing import List

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libgdbnerf_b200.so")
STAMP = LIB + ".stamp"
SOURCES = ["gdb_costvolume.cu", "gdb_probhead_tma.cu", "gdb_sampling.cu", "gdb_prepare.cu", "gdb_render.cu", "gdb_render_tc.cu", "gdb_render_tc2.cu", "gdb_render_tc3.cu", "gdb_render_tc4.cu", "gdb_glue.cu", "gdb_costvolume_bwd.cu", "gdb_render_bwd.cu", "gdb_coarse.cu", "gdb_optim.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the gdb_nerf_b200 CUDA library cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "gdb_nerf_b200.h")]
    for f in files:
        path = os.path.join(CSRC, f)
        if os.path.isfile(path):
            h.update(f.encode())
            with open(path, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objs: List[str] = []
    procs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(PKG, "build", src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(f"--- {src}\n{out}\n")
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed; see output above")
    link = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
    subprocess.run(link, check=True)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
