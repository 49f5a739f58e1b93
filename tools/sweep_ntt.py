#!/usr/bin/env python3
"""NTT tuning sweep: compile-time variants of csrc/ntt.cu x run-time split / column-group settings.

    python tools/sweep_ntt.py build            # here (no GPU): lib/libr0b200_ntt_<variant>.so for every variant
    python tools/sweep_ntt.py run [--lg 20] [--cols 64] > gpurun_out/ntt_sweep.log    # on the GPU box

Only ntt.cu is recompiled per variant; every other object comes from the main build directory (risc0_b200/build).
Each (variant, env) point runs tools/bench_ntt.py in a fresh process: iNTT+zk at 2^lg, expand+NTT 2^lg -> 2^(lg+2) and
bit-reverse, per-op device time from CUDA events on the launching stream.
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
PKG = os.path.join(ROOT, "risc0_b200")
NVCC = "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]

# Sweep 1 (profiles/r1_ntt_sweep1.log) fixed: row-table twiddles + lazy operands on, contiguous pass 2^12, no L2 column
# groups. Sweep 2 (r1_ntt_sweep2.log): strided-pass tile width x block size: 8 columns x 256 threads stays.
# Sweep 3 (r1_ntt_sweep3.log): Shoup twiddle products / butterfly adds as IMAD: no effect (ptxas already emits IMAD.IADD).
# Sweep 4: the opposite direction - keep adds on the alu pipe (three-input IADD3), with and without Shoup products.
# Sweep 5 (r2_ntt_sweep5.log): additive Montgomery reduction in the twiddle products.
VARIANTS = {
    "base": [],
    "additive": ["-DR0_NTT_ADDITIVE=1"],
    "additive_iadd3": ["-DR0_NTT_ADDITIVE=1", "-DR0_NTT_IADD3=3"],
}

# run-time points tried for every variant (the first one is the library default)
ENVS = [
    {},
]


def lib_path(name):
    return os.path.join(PKG, "lib", "libr0b200_ntt_%s.so" % name)


def build():
    from risc0_b200 import build as b
    b.build()
    # exactly the objects the main library links (the build directory may hold stale ones), minus ntt.o
    others = [os.path.join(b.OBJ, os.path.relpath(src, b.CSRC).replace(os.sep, "_")[:-3] + ".o") for src in b.sources()]
    others = [o for o in others if not o.endswith(os.sep + "ntt.o")] + [os.path.join(b.OBJ, "embed_cubins.o")]
    os.makedirs(os.path.join(PKG, "build_ntt"), exist_ok=True)
    procs = []
    for name, defs in VARIANTS.items():
        obj = os.path.join(PKG, "build_ntt", "ntt_%s.o" % name)
        cmd = [NVCC] + FLAGS + defs + ["-I", os.path.join(PKG, "csrc"), "-c", os.path.join(PKG, "csrc", "ntt.cu"), "-o", obj]
        procs.append((name, obj, subprocess.Popen(cmd)))
    for name, obj, p in procs:
        if p.wait() != 0:
            raise SystemExit("compile failed: " + name)
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path(name), obj] + others)
        print("built", lib_path(name))


def run(lg, cols, variants):
    for name in variants:
        for env in ENVS:
            e = dict(os.environ, R0B200_LIB=lib_path(name), **env)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_ntt.py"), "--lg", str(lg), "--cols", str(cols),
                                "--iters", "10"], env=e, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else None
            rec = {"variant": name, "env": env}
            if line:
                d = json.loads(line)
                rec.update({k: v["ms"] for k, v in d.items() if isinstance(v, dict)})
                rec["shape"] = d["shape"]
            else:
                rec["error"] = (r.stderr or "")[-300:]
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    if sys.argv[1] == "build":
        build()
    else:
        import argparse
        ap = argparse.ArgumentParser()
        ap.add_argument("cmd")
        ap.add_argument("--lg", type=int, default=20)
        ap.add_argument("--cols", type=int, default=64)
        ap.add_argument("--variants", default=",".join(VARIANTS))
        a = ap.parse_args()
        run(a.lg, a.cols, a.variants.split(","))
