"""Ahead-of-time build of libfvvdp_b200.so (nvcc, sm_100a only, in-tree so that it travels with gpurun)."""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc", "fvvdp_b200.cu")
DEPS = [SRC, os.path.join(PKG, "csrc", "fvvdp_kernels.cuh"), os.path.join(os.path.dirname(PKG), "include", "fvvdp_b200.h")]
LIB = os.path.join(PKG, "_lib", "libfvvdp_b200.so")


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: libfvvdp_b200.so cannot be built (there is no CPU fallback)")


def is_stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build_native(force=False, verbose=False):
    """Compile csrc/fvvdp_b200.cu -> _lib/libfvvdp_b200.so for sm_100a.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3", "-o", LIB + ".tmp", SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        sys.stderr.write(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_native(force=True, verbose="-v" in sys.argv))
