#!/usr/bin/env python
"""Build libb200dsp.so (hand-written CUDA for sm_100a) in-tree with nvcc.

    python scikit-dsp-comm_b200/build.py [--force] [--verbose]

Output: scikit-dsp-comm_b200/sk_dsp_comm_b200/_lib/libb200dsp.so (git-ignored; it travels to
the GPU box with the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "sk_dsp_comm_b200", "_lib")
OBJ_DIR = os.path.join(HERE, "build")
SO = os.path.join(OUT_DIR, "libb200dsp.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newest_dep():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    dep_t = _newest_dep()
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= dep_t:
        return SO
    ccbin = ["-ccbin", HOST_CXX] if HOST_CXX else []

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        path = os.path.join(CSRC, src)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= dep_t):
            return obj
        cmd = [NVCC] + ccbin + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for " + src)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [NVCC] + ccbin + ["-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO] + objs
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
