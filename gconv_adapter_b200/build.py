"""Build libgca in-tree with nvcc for sm_100a:  python -m gconv_adapter_b200.build [--force] [-v]

Every translation unit under csrc/ is compiled separately (in parallel) and linked into
``lib/libgca.<hash>.so`` where <hash> is the SHA-256 of the sources, the public header and the nvcc
flags.  ``_cabi.load()`` computes the same hash and loads exactly that file, so a stale binary can
never be picked up for changed sources: it is simply "not built" (and there is no CPU fallback).
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _hash_inputs() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*"))) + [os.path.join(ROOT, "include", "gca.h")]


def source_hash() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + os.environ.get("GCA_EXTRA_NVCC_FLAGS", "").split()).encode())
    for path in _hash_inputs():
        if os.path.isfile(path):
            h.update(os.path.basename(path).encode())
            with open(path, "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]


def lib_path() -> str:
    return os.environ.get("GCA_BUILD_OUT") or os.path.join(LIB_DIR, f"libgca.{source_hash()}.so")


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libgca")


def build(force: bool = False, verbose: bool = False) -> str:
    out = lib_path()
    if not force and os.path.isfile(out):
        return out
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(ROOT, "build", "obj", source_hash())
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = nvcc_path()
    flags = NVCC_FLAGS + ["-I" + os.path.join(ROOT, "include"), "-Xcompiler", "-fPIC"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    flags += os.environ.get("GCA_EXTRA_NVCC_FLAGS", "").split()

    def compile_one(src: str) -> tuple[str, str]:
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        res = subprocess.run([nvcc] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {os.path.basename(src)}:\n{res.stdout}{res.stderr}")
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    if verbose:
        for _, log in results:
            print(log)
    tmp = out + ".tmp"
    res = subprocess.run([nvcc, "--shared", "-o", tmp] + [o for o, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, out)
    for old in glob.glob(os.path.join(LIB_DIR, "libgca*.so")):      # one binary per tree: drop builds of older sources
        if os.path.abspath(old) != os.path.abspath(out):
            os.remove(old)
    for old in glob.glob(os.path.join(ROOT, "build", "obj", "*")):  # ... and their objects
        if os.path.abspath(old) != os.path.abspath(obj_dir):
            shutil.rmtree(old, ignore_errors=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
