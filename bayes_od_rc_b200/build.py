"""Builds the CUDA shared library (bayes_od_rc_b200/lib/libbayesod.so) in-tree
with nvcc for sm_100a.  Called by __graft_entry__.build(); also usable as
``python -m bayes_od_rc_b200.build``.

Two flag sets: the moments kernel (k1) may contract FMAs (its softmax output is
tolerance-checked); every other translation unit is compiled with -fmad=false
because it implements the bit-exact arithmetic contract."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libbayesod.so")
OBJDIR = os.path.join(LIBDIR, "obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math",
          "--ftz=false", "--prec-div=true", "--prec-sqrt=true"]
UNITS = {
    "k1_moments.cu": [],                      # FMA allowed (explicit __fmaf_rn / __expf only in the softmax)
    "k2_posterior.cu": ["-fmad=false"],
    "k3_softnms.cu": ["-fmad=false"],
    "k4_fusion.cu": ["-fmad=false"],
    "kv_validation.cu": ["-fmad=false"],
    "bod_api.cu": ["-fmad=false"],
    "bod_io.cu": [],                          # host-only: npy / json / txt writers
    "kp_pdq.cu": [],                          # PDQ heat maps and loss sums: binary64 CDF, tolerance-checked
    "ku_uncertainty.cu": ["-fmad=false"],     # entropies + MUE curve (binary64 like the reference's numpy)
}
HEADERS = ["bod_common.cuh", "bod_kernels.h", os.path.join("..", "..", "include", "bayesod.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    nvcc = _nvcc()
    for unit, extra in UNITS.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJDIR, unit.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            extra = extra + (os.environ.get("BOD_EXTRA_NVCC_" + unit.split(".")[0].upper(), "").split())
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if verbose or r.returncode:
                print(r.stdout, file=sys.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed for {unit}")
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            print(r.stdout, file=sys.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
