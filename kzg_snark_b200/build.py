"""Build libkzgpu.so (sm_100a only) in-tree with nvcc.  No torch, no JIT cache.

    python -m kzg_snark_b200.build          # or: python kzg_snark_b200/build.py

The .so stays inside the package directory (git-ignored) so it travels with the tree.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libkzgpu.so")
OUT_BENCH = os.path.join(HERE, "libkzgpu_bench.so")          # microbenchmarks: measurement tooling, separate from the product library
SOURCES = ["context.cu", "ntt.cu", "msm.cu", "poly.cu", "plonk.cu"]
BENCH_SOURCES = ["microbench.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _needs_rebuild():
    if not os.path.exists(OUT) or not os.path.exists(OUT_BENCH):
        return True
    t = os.path.getmtime(OUT)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh")) and os.path.getmtime(os.path.join(root, f)) > t:
                return True
    inc = os.path.join(os.path.dirname(HERE), "include")
    return max(os.path.getmtime(os.path.join(inc, f)) for f in os.listdir(inc)) > t


def build(force=False, verbose=False):
    if not force and not _needs_rebuild():
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    hdr_t = max([os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(".cuh")] +
                [os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", f)) for f in ("kzgpu.h", "kzgpu_bench.h")])

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if (not force and not verbose and os.path.exists(obj)
                and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(os.path.join(CSRC, src)))):
            return obj                                   # object is newer than its source and every header
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(compile_one, SOURCES + BENCH_SOURCES))
    for out, parts in ((OUT, objs[:len(SOURCES)]), (OUT_BENCH, objs[len(SOURCES):])):
        cmd = [NVCC, "-shared", "-o", out, *parts, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
