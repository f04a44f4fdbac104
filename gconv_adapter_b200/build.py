"""Build libgca.so in-tree with nvcc for sm_100a:  python -m gconv_adapter_b200.build"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = ["gca_lib.cu", "gca_graph.cu", "gca_kernels.cu", "gca_tc_project.cu"]
OUT = os.environ.get("GCA_BUILD_OUT") or os.path.join(HERE, "lib", "libgca.so")


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libgca.so")


def needs_build() -> bool:
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    deps.append(os.path.join(ROOT, "include", "gca.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-I" + os.path.join(ROOT, "include"), "--shared", "-Xcompiler", "-fPIC", "-o", OUT]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("GCA_EXTRA_NVCC_FLAGS", "").split()
    cmd += [os.path.join(HERE, "csrc", s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
