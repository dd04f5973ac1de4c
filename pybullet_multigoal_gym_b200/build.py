"""Builds libpmg.so (the CUDA kernels + C-ABI) in-tree for sm_100a with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libpmg.so")
SOURCES = ["pmg_capi.cu"]


def _deps():
    """Every source the library is built from: all of csrc/ and include/ (a stale libpmg.so must never be measured)."""
    inc = os.path.join(HERE, "..", "include")
    return ([os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
            + [os.path.join(inc, f) for f in sorted(os.listdir(inc)) if f.endswith(".h")] + [os.path.abspath(__file__)])


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "-prec-div=false", "-prec-sqrt=false", "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if verbose or res.returncode != 0:
        sys.stderr.write(log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpmg.so (see build.log)")
    return SO


def build_timing():
    """Development build with per-phase cycle counters in the cooperative kernels (tools/coop_timing.py):
    libpmg_timing.so, selected with PMG_LIBRARY; never loaded by default."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.path.join(HERE, "libpmg_timing.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-DPMG_COOP_TIMING", "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libpmg_timing.so")
    return out


if __name__ == "__main__":
    if "--timing" in sys.argv:
        print(build_timing())
    else:
        build(force=True, verbose="-v" in sys.argv)
        print(SO)
