"""Build libmarex_b200.so in-tree with nvcc for sm_100a (no torch extension machinery:
the library is a plain C-ABI shared object, loaded with ctypes)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmarex_b200.so")
SOURCES = ["anomaly.cu", "shift_daily.cu", "thresholds.cu", "pool_band.cu", "compare.cu"]
HEADERS = ["common.cuh", "tma.cuh", os.path.join("..", "..", "include", "marex_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [
        _nvcc(),
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-O3", "-lineinfo", "-std=c++17",
        "--shared", "-Xcompiler", "-fPIC",
        "-Xptxas", "-v" if verbose else "-O3",
        "-o", LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libmarex_b200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
