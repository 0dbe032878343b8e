"""Build the sm_100a shared library in-tree (``wab_gym_b200/libwab_b200.so``) with nvcc.

The library is plain CUDA C++ behind the C ABI of ``include/wab_b200.h``; it links only cudart, so a
direct ``nvcc -shared`` is the whole build (no torch headers, no JIT cache). ``python -m
wab_gym_b200.build`` or ``__graft_entry__.build()`` runs it; nvcc cross-compiles without a GPU.
"""
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libwab_b200.so")
SOURCES = [os.path.join(CSRC, "wab_kernels.cu")]


def headers():
    """Every header the translation unit can see: an edit to any of them makes the library stale."""
    inc = os.path.join(os.path.dirname(PKG_DIR), "include")
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  glob.glob(os.path.join(inc, "*.h")))


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",            # fp64 food must follow the reference's op order exactly (no contraction)
    "-Xcompiler", "-fPIC", "-shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC=/path/to/nvcc")


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > built for p in SOURCES + headers())


def build(force=False, verbose=False):
    """Compile if the library is missing or older than its sources. Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout)
    if verbose:
        print(proc.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
