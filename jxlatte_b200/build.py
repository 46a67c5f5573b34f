"""Builds jxlatte_b200/libjxlb200.so for sm_100a with nvcc (in-tree, so the .so travels with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.environ.get("JXLB_SO") or os.path.join(HERE, "libjxlb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++",
    "-Xptxas", "-v" if os.environ.get("JXLB_PTXAS_V") else "-O3",
] + os.environ.get("JXLB_EXTRA_FLAGS", "").split()


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(HERE, "..", "include", "jxlb200.h")]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [NVCC] + FLAGS + ["-o", SO, os.path.join(CSRC, "jxlb200.cu"), "-lcudart", "-ldl"]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libjxlb200.so")
    if verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(SO)
