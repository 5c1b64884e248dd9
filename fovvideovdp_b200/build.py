"""Ahead-of-time build of libfvvdp_b200.so (nvcc, sm_100a only, in-tree so that it travels with gpurun).
The translation units are compiled in parallel and linked into one shared library."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
HEADERS = [os.path.join(CSRC, h) for h in ("fvvdp_common.cuh", "fvvdp_kernels.cuh", "fvvdp_fused.cuh", "fvvdp_fused_launch.h", "fvvdp_ws.cuh", "fvvdp_ws_geometry.h")] + \
    [os.path.join(os.path.dirname(PKG), "include", "fvvdp_b200.h")]
# (object name, source, extra defines)
UNITS = [("fvvdp_b200", "fvvdp_b200.cu", []), ("fused_dispatch", "fvvdp_fused_dispatch.cu", [])] + \
    [(f"fused_{k}_{v}", "fvvdp_fused_inst.cu", [f"-DFUSED_KIND={k}", f"-DFUSED_VIDEO={v}"]) for k in (0, 2, 3) for v in range(3)] + \
    [("fused_2_3", "fvvdp_fused_inst.cu", ["-DFUSED_KIND=2", "-DFUSED_VIDEO=3"])] + \
    [(f"ws{'16' if t else ''}_{k}", "fvvdp_ws_inst.cu", [f"-DWS_KIND={k}", f"-DWS_TAPS16={t}"]) for k in (2, 3) for t in (0, 1)]
DEPS = HEADERS + [os.path.join(CSRC, u[1]) for u in UNITS]
OBJ_DIR = os.path.join(PKG, "_lib", "obj")
LIB = os.environ.get("FVVDP_B200_LIB") or os.path.join(PKG, "_lib", "libfvvdp_b200.so")  # override: experiment builds
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found: libfvvdp_b200.so cannot be built (there is no CPU fallback)")


def is_stale():
    if os.environ.get("FVVDP_B200_LIB"):
        return False
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def _compile(unit, verbose):
    name, src, defs = unit
    obj = os.path.join(OBJ_DIR, name + ".o")
    srcp = os.path.join(CSRC, src)
    if os.path.isfile(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in HEADERS + [srcp]):
        return obj, ""
    cmd = [nvcc_path()] + ARCH + FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return obj, r.stderr


def build_native(force=False, verbose=False):
    """Compile csrc/*.cu -> _lib/libfvvdp_b200.so for sm_100a.  Returns the library path."""
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as ex:
        res = list(ex.map(lambda u: _compile(u, verbose), UNITS))
    cmd = [nvcc_path()] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB + ".tmp"] + [r[0] for r in res]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        for u, (_, log) in zip(UNITS, res):
            sys.stderr.write(f"==== {u[0]} ====\n{log}")
    return LIB


if __name__ == "__main__":
    print(build_native(force="-f" in sys.argv, verbose="-v" in sys.argv))
