"""Build libmarex_b200.so in-tree with nvcc for sm_100a (no torch extension machinery:
the library is a plain C-ABI shared object, loaded with ctypes)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmarex_b200.so")
SOURCES = ["anomaly.cu", "shift_daily.cu", "thresholds.cu", "pool_band.cu", "compare.cu", "morph.cu"]
HEADERS = ["common.cuh", "tma.cuh", "digitize.cuh", "exact_queue.cuh", "morph_core.cuh", os.path.join("..", "..", "include", "marex_b200.h")]


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
    """Compile every .cu to an object file in parallel, then link the shared library."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
             "-Xptxas", "-v" if verbose else "-O3"]  # fmt: skip
    flags += os.environ.get("MAREX_NVCC_FLAGS", "").split()  # experiments (tools/gpu_*.sh build variants on the GPU box)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        res = subprocess.run([nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        return src, obj, res

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for src, _obj, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed compiling {src}")
        if verbose:
            sys.stderr.write(res.stdout + res.stderr)
    link = subprocess.run([nvcc, "--shared", "-o", LIB] + [o for _s, o, _r in results], capture_output=True, text=True)
    if link.returncode != 0:
        sys.stderr.write(link.stdout + link.stderr)
        raise RuntimeError("nvcc failed linking libmarex_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
