"""Build libcvcl_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python multimodal-baby_b200/build.py        # or: from multimodal_baby_b200 import build

nvcc cross-compiles for sm_100a without a GPU.  The .so links only cudart (static) -- no
libtorch, no libcuda (the TMA descriptor encoder is fetched through cudaGetDriverEntryPoint).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libcvcl_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
              "--shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libcvcl_b200.so")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                  if f.endswith((".cu", ".cuh", ".h")))


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    hdr = os.path.join(os.path.dirname(HERE), "include", "cvcl_b200.h")
    return any(os.path.getmtime(s) > t for s in sources() + [hdr])


def build_library(force=False, verbose=False):
    """Compile csrc/cvcl_b200.cu -> lib/libcvcl_b200.so.  Returns the library path.
    Safe under torchrun: ranks serialise on a file lock, nvcc writes a temporary file and the finished
    library is renamed into place (a rank never dlopens a half-written .so)."""
    if not force and not is_stale():
        return LIB_PATH
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():         # another rank built it while we waited
                return LIB_PATH
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                  ["-diag-suppress", "39", "-o", tmp, os.path.join(CSRC, "cvcl_b200.cu")]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def build_selftest():
    out = os.path.join(LIB_DIR, "selftest_gemm")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
           "-o", out, os.path.join(CSRC, "selftest_gemm.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
