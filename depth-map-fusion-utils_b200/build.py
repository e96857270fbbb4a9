"""Builds libdmf_b200.so (the C-ABI CUDA library) in-tree for sm_100a.

    python depth-map-fusion-utils_b200/build.py [--force] [--verbose]

-fmad=false: the kernels must reproduce the reference's separate IEEE mul/add results bit for bit, so implicit
FMA contraction is off and fused ops are spelled explicitly where they are provably safe.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("DMF_B200_OUT") or os.path.join(HERE, "libdmf_b200.so")       # DMF_B200_OUT: build a variant next to the default
SOURCES = ["dmf_b200.cu"]
DEPS = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, "..", "include", "dmf_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared",
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdmf_b200.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    extra = os.environ.get("DMF_NVCC_EXTRA", "").split()          # e.g. "-DDMF_SKIP_WARPS_X=4" for tile-shape experiments
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); nvcc should use the distro g++ on PATH
    env.pop("CXX", None); env.pop("CC", None)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return OUT


def build_checked(force: bool = False) -> str:
    """libdmf_b200_checked.so: the same sources with -DDMF_CHECKED (every computed grid index bounds-checked and counted,
    dmf_device.cuh); tests/test_fuzz_gpu.py runs the fuzz against it.  compute-sanitizer is not available on the GPU pool."""
    global OUT
    out = os.path.join(HERE, "libdmf_b200_checked.so")
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in DEPS):
        return out
    saved, saved_env = OUT, os.environ.get("DMF_NVCC_EXTRA")
    OUT = out
    os.environ["DMF_NVCC_EXTRA"] = ((saved_env or "") + " -DDMF_CHECKED").strip()
    try:
        return build(force=True)
    finally:
        OUT = saved
        if saved_env is None:
            os.environ.pop("DMF_NVCC_EXTRA", None)
        else:
            os.environ["DMF_NVCC_EXTRA"] = saved_env


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
