"""Build libkvq.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python kindergarten-vq-vae_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB = os.path.join(PKG_DIR, "libkvq.so")
STAMP = os.path.join(PKG_DIR, "build", "libkvq.stamp")
SOURCES = ["api.cu", "bandwidth_kernels.cu", "backward.cu", "search_fp32.cu", "search_tf32.cu", "aux.cu", "recon.cu", "gumbel.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libkvq cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(PKG_DIR, "..", "include", "kvq.h"))
    files.append(os.path.abspath(__file__))
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources changed; returns the path of the shared library."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    os.makedirs(os.path.dirname(STAMP), exist_ok=True)
    flags = [f for f in FLAGS if not f.startswith("--use_fast_math")]
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(PKG_DIR, "build", src.replace(".cu", ".o"))
        cmd = [_nvcc(), *ARCH, *flags, "-Xcompiler", "-fPIC", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            print(f"[kvq build] {src} FAILED:\n{out}", file=sys.stderr)
        elif verbose or out.strip():
            print(f"[kvq build] {src}:\n{out}")
    if failed:
        raise RuntimeError("nvcc failed building libkvq")
    link = [_nvcc(), *ARCH, "-shared", "-o", LIB, *objs]
    subprocess.run(link, check=True)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
