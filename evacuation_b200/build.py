"""In-tree build of the CUDA library (sm_100a only).  `python -m evacuation_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("EVAC_B200_LIB") or os.path.join(PKG_DIR, "libevac_b200.so")  # override: kernel-variant experiments
SOURCES = [os.path.join(CSRC, "evac_abi.cu"), os.path.join(CSRC, "evac_policy.cu")]
HEADERS = [os.path.join(CSRC, "evac_kernels.cuh"), os.path.join(CSRC, "evac_warp.cuh"), os.path.join(CSRC, "philox.cuh"),
           os.path.join(CSRC, "evac_policy.cuh"), os.path.join(os.path.dirname(PKG_DIR), "include", "evac_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; evacuation_b200 has no CPU fallback and needs the CUDA toolkit to build")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libevac_b200.so next to this file (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = find_nvcc()
    objs, procs = [], []
    for src in SOURCES:  # one nvcc per translation unit, in parallel
        obj = os.path.join(CSRC, os.path.splitext(os.path.basename(src))[0] + ".o")
        obj = obj[:-2] + os.environ.get("EVAC_B200_OBJ_TAG", "") + ".o"
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("EVAC_B200_NVCC_EXTRA", "").split(), *(["-Xptxas=-v"] if verbose else []), "-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        objs.append(obj)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for cmd, pr in procs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{out}\n{err}")
        if verbose:
            print(err, file=sys.stderr)
    res = subprocess.run([nvcc, "-shared", "-o", LIB_PATH, *objs], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
