"""In-tree build of libr0b200.so (hand-written sm_100a CUDA + C ABI) with plain nvcc.

    python -m risc0_b200.build            # incremental
    python -m risc0_b200.build --force

The library lands in risc0_b200/lib/ (git-ignored, but shipped to the GPU box with the snapshot). There is no
CPU fallback: importing risc0_b200 on a machine where this library is missing raises.
"""
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libr0b200.so")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "gen", "*.cu")))


def headers():
    return (glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
            glob.glob(os.path.join(CSRC, "tables", "*.h")) + glob.glob(os.path.join(CSRC, "gen", "*.h")) +
            glob.glob(os.path.join(CSRC, "gen", "*.cuh")) + [os.path.join(HERE, "..", "include", "r0b200.h")])


def _compile(src, obj, verbose):
    cmd = [NVCC] + FLAGS + ["-I", CSRC, "-c", src, "-o", obj]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return src, r.stderr


def build(force=False, verbose=False, jobs=None):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in headers())
    todo, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.relpath(src, CSRC).replace(os.sep, "_")[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            todo.append((src, obj))
    if todo:
        with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            for src, log in ex.map(lambda so: _compile(so[0], so[1], verbose), todo):
                if verbose:
                    sys.stderr.write("== %s\n%s" % (os.path.basename(src), log))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
