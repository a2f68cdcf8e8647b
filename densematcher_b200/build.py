"""Builds densematcher_b200/libdm_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, no JIT cache).

    python -m densematcher_b200.build [--force] [--verbose]

The shared library is a plain C-ABI object (include/dm_b200.h): no torch, no pybind.  nvcc
cross-compiles on a machine without a GPU; the resulting .so travels to the GPU box with the tree.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(PKG, "libdm_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libdm_b200.so")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "dm_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src)
            and os.path.getmtime(obj) >= _deps_mtime()):
        return obj, ""
    with open(src) as f:                      # per-file flags: a first line "// dm-nvcc-flags: ..."
        first = f.readline()
    extra = first.split("dm-nvcc-flags:", 1)[1].split() if "dm-nvcc-flags:" in first else []
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    log = r.stderr
    with open(obj[:-2] + ".ptxas.log", "w") as f:
        f.write(log)
    if verbose:
        print(log)
    return obj, log


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the C-ABI shared library.  Returns its path."""
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = [o for o, _ in ex.map(lambda s: _compile(s, force, verbose), srcs)]
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
               "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
