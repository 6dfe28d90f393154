"""Build libtrl_b200.so (hand-written sm_100a kernels + the C ABI of include/trl.h) in-tree.

    python -m tetris_reinforcement_learning_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with gpurun snapshots.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "libtrl_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(PKG), "include", "trl.h")]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(f) > t for f in _deps())


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a into one shared library."""
    if not force and not needs_build():
        return SO
    cmd = [NVCC] + FLAGS + os.environ.get("TRL_NVCC_EXTRA", "").split() + ["-o", SO] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(PKG, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    if verbose:
        print(log)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", SO)
