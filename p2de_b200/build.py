"""Builds libp2de_b200.so in-tree with nvcc for sm_100a (no JIT cache, no pip install)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libp2de_b200.so")
SOURCES = ["capi.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              # ptxas compiles the kernels of the one translation unit in parallel.  NOT nvcc's own `-split-compile 0`: that also
              # splits the NVVM optimiser (cicc), whose output for this file is then not reproducible -- the same command gives one
              # of two different PTX files from run to run (different inlining / unswitching, 24..144 B of stack in the stage
              # kernels); with cicc unsplit the PTX and the SASS of every kernel are identical across builds.
              "-Xptxas", "-split-compile=0",
              "-diag-suppress", "177",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "p2de_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_extension(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return SO
    extra = os.environ.get("P2DE_NVCC_EXTRA", "").split()   # A/B builds: -D switches of the kernel headers
    so = os.environ.get("P2DE_B200_LIB", SO) if extra else SO
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", so] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CXX", None), env.pop("CC", None)
    res = subprocess.run(cmd, cwd=CSRC, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return so


if __name__ == "__main__":
    import sys
    print(build_extension(force=True, verbose="-v" in sys.argv))
