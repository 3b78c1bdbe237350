"""Build libbpx.so in-tree for sm_100a (B200).  `python build.py [--force] [--verbose]`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get("BPX_BUILD_OUT", os.path.join(HERE, "libbpx.so"))  # BPX_BUILD_OUT: debug builds next to the product
SOURCES = ["bpx_api.cu"]
DEPS = [f for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".h"))] + [os.path.join("..", "..", "include", "bpx.h")]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(HERE, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [
        nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
        "-Xcompiler", "-fPIC", "-shared", "-o", OUT,
    ] + os.environ.get("NVCC_EXTRA", "").split() + [os.path.join(HERE, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libbpx.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
