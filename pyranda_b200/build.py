"""Build libparcop_b200.so in-tree for sm_100a (`python -m pyranda_b200.build`).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the repo
snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libparcop_b200.so")
SOURCES = ["tables.cpp", "kernels.cu", "api.cu", "sweeps_d1.cu", "sweeps_r3.cu", "sweeps_r4.cu", "sweeps_r4v.cu"]
DEPS = SOURCES + ["kernels.cuh", "sweeps.cuh", "ring.cuh", "tma.cuh", "tables.hpp", os.path.join("..", "..", "include", "parcop_b200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]
OBJ_DIR = os.path.join(HERE, "build")


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    header_time = max(os.path.getmtime(os.path.join(CSRC, d)) for d in DEPS if not d.endswith((".cu", ".cpp")))

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        t_src = max(os.path.getmtime(os.path.join(CSRC, src)), header_time)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < t_src:
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            subprocess.check_call(cmd, cwd=CSRC)
        return obj

    # the stencil families are separate translation units so they compile in parallel
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([nvcc, "-shared", "-o", OUT] + objs, cwd=CSRC)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
