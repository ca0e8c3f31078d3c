"""Builds libnpp_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

`python -m npp_b200.build` or `npp_b200.build.build_lib()`; `__graft_entry__.build()` calls this.
nvcc cross-compiles for sm_100a without a GPU.  Object files are cached by source mtime.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libnpp_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
    "-I", os.path.join(ROOT, "include"),
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(spath), _headers_mtime()):
        return obj
    cmd = [NVCC] + FLAGS + ["-c", spath, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def build_lib(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


def build_test_binaries(verbose=False):
    """Standalone device tests under tests/csrc (linked against libnpp_b200.so)."""
    lib = build_lib(verbose)
    out_dir = os.path.join(ROOT, "tests", "csrc", "_bin")
    os.makedirs(out_dir, exist_ok=True)
    outs = []
    tdir = os.path.join(ROOT, "tests", "csrc")
    for f in sorted(os.listdir(tdir)):
        if not f.endswith(".cu"):
            continue
        out = os.path.join(out_dir, f[:-3])
        src = os.path.join(tdir, f)
        if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(lib)):
            outs.append(out)
            continue
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", src, "-o", out,
               "-L", HERE, "-lnpp_b200", "-ldl", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../../npp_b200"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (f, r.stdout, r.stderr))
        outs.append(out)
    return outs


if __name__ == "__main__":
    print(build_lib(verbose=True, force="--force" in sys.argv))
    if "--tests" in sys.argv:
        print(build_test_binaries(verbose=True))
