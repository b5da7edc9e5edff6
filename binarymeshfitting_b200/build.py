"""In-tree build of libbmf_b200.so (hand-written sm_100a CUDA + the C ABI) with nvcc.

No torch, no JIT cache: the .so lands next to the sources so it travels with the repo snapshot.
-fmad=false: the kernels place every fused multiply-add explicitly (__fmaf_rn) and nothing else may be
contracted -- bit-exact vertex positions and noise depend on it.  -lineinfo keeps ncu's source page usable.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libbmf_b200.so")
SOURCES = ["bmf_b200.cu"]
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [os.path.join("..", "..", "include", "bmf_b200.h")]
NVCC = os.environ.get("BMF_NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS) or os.path.getmtime(__file__) > t


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [NVCC] + FLAGS + ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libbmf_b200.so")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(r.stderr)
    if verbose:
        sys.stderr.write(r.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
