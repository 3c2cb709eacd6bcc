"""Builds gym_art_b200/libquadsim.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m gym_art_b200._build [--force] [--verbose]

The translation units compile in parallel; the shared object is git-ignored but travels to the GPU box
with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libquadsim.so")
FAST = ["qs_fast_%s_%s.cu" % (c, n) for c in ("raw", "omega", "mellinger", "shifted", "velyaw") for n in ("philox", "inject")]
GENERIC = ["qs_tma_generic_%s_%s.cu" % (f, n) for f in ("all", "drag", "noise", "misc", "info", "dragnoise") for n in ("philox", "inject")]
SOURCES = ["qs_kernels_f32.cu", "qs_kernels_f64.cu", "quadsim_abi.cu"] + GENERIC + FAST
DEPS = ["qs_device.cuh", "qs_kernels.inl", "qs_fast.inl", "qs_tma_launch.inl", "qs_tma_generic.inl", os.path.join("..", "..", "include", "quadsim.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# no --use_fast_math: parity with the float64 reference is part of the contract (fast intrinsics are used
# explicitly, and only inside the Philox Box-Muller transform)
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _stale(lib):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    files = [os.path.join(CSRC, f) for f in SOURCES + DEPS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(f) > t for f in files)


# experiment / debug variants of the library: name -> (extra nvcc defines, output path).  They are built on request only
# (python -m gym_art_b200._build --variant bounds) and loaded through QUADSIM_LIB; the product is the plain library.
VARIANTS = {
    "bounds": (["-DQS_DEBUG_BOUNDS"], os.path.join(HERE, "_exp", "libquadsim_bounds.so")),     # every derived index checked
    "timing": (["-DQS_PHASE_TIMING"], os.path.join(HERE, "_exp", "libquadsim_timing.so")),     # %globaltimer stamps per warp (tools/phase_timing.py)
}


def build(force=False, verbose=False, variant=None):
    """Compile if sources are newer than the library.  Returns the library path."""
    defines, lib = VARIANTS[variant] if variant else ([], LIB)
    if not force and not _stale(lib):
        return lib
    bdir = BUILD if not variant else os.path.join(BUILD, variant)
    os.makedirs(bdir, exist_ok=True)
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(bdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + defines + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-8000:]))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [nvcc] + ARCH + ["-shared", "-o", lib + ".tmp"] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    os.replace(lib + ".tmp", lib)
    return lib


if __name__ == "__main__":
    v = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, variant=v)
    print(path)
