"""Build libgridinterp.so (CUDA, sm_100a) in-tree with nvcc.

    python -m grid_interpolation_b200.build [--force]

nvcc cross-compiles without a GPU.  Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no other architectures, no PTX fallback
  -fmad=false                               no FMA contraction: every index decision and value follows the
                                            reference's rounding sequence (SURVEY section 7 "hard parts")
  -lineinfo                                 ncu source view
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgridinterp.so")
SOURCES = ["api.cu", "bilinear.cu", "barycentric.cu", "esrg.cu", "kdtree.cu", "mesh.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3", "-Xptxas", "-v"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libgridinterp.so cannot be built")
    return nvcc


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "gridinterp.h"))
    return hdrs


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    srcs = [os.path.join(CSRC, s) for s in SOURCES] + _deps() + [os.path.abspath(__file__)]
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False, defines=(), out_dir: str | None = None) -> str:
    """Build the library.  ``defines`` / ``out_dir``: an A/B variant (extra -D flags, objects and .so under
    ``out_dir``; load it with GI_LIBRARY=<out_dir>/libgridinterp.so)."""
    if out_dir:
        return _build(tuple(defines), os.path.join(out_dir, "build"), os.path.join(out_dir, "libgridinterp.so"), verbose)
    if not force and not needs_build():
        return LIB
    return _build((), BUILD, LIB, verbose)


def _build(defines, BUILD, LIB, verbose) -> str:
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []

    def compile_one(src):
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc, *ccbin, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            print(log)
    objs = [o for o, _ in results]
    cmd = [nvcc, *ccbin, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB + ".tmp", *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    # python -m grid_interpolation_b200.build [--force] [-v] [--variant DIR -DNAME=VALUE ...]
    if "--variant" in sys.argv:
        d = sys.argv[sys.argv.index("--variant") + 1]
        print(build(defines=[a[2:] for a in sys.argv if a.startswith("-D")], out_dir=d, verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
