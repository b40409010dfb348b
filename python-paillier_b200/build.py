"""Build the CUDA engine in-tree:  python-paillier_b200/libpaillier_b200.so  (sm_100a only).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repo snapshot.  Rebuilds only when a source is newer than the library.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpaillier_b200.so")
SOURCES = ["pai_engine.cu"]
HEADERS = ["pai_core.cuh", "pai_kernels.cuh", "pai_digit.cuh", "pai_cta.cuh", "pai_coop.cuh", "pai_tc.cuh", "pai_rng.cuh", "pai_radix.cuh", "pai_rt.h", os.path.join("..", "..", "include", "paillier_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        sys.stderr.write(log[-4000:])
        raise RuntimeError("nvcc failed (see %s/build.log)" % HERE)
    if verbose:
        print(log[-2000:])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
