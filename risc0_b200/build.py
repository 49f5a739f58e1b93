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
VARIANT = os.environ.get("R0B200_VARIANT", "")   # experiments: separate object dir + library name
OBJ = os.path.join(HERE, "build" + ("_" + VARIANT if VARIANT else ""))
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libr0b200%s.so" % ("_" + VARIANT if VARIANT else ""))
PTXAS_OPT = os.environ.get("R0B200_PTXAS_OPT", "-O3")
PTXAS_EXTRA = os.environ.get("R0B200_PTXAS_EXTRA", "").split()

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "gen", "*.cu")))


def headers():
    return (glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
            glob.glob(os.path.join(CSRC, "tables", "*.h")) + glob.glob(os.path.join(CSRC, "gen", "*.h")) +
            glob.glob(os.path.join(CSRC, "gen", "*.cuh")) + [os.path.join(HERE, "..", "include", "r0b200.h")])


def ptx_sources():
    return sorted(glob.glob(os.path.join(CSRC, "gen", "*.ptx")))


def generate(force=False):
    """(re)create csrc/gen/* from the committed circuit IR when missing or stale (tools/gen_eval_check.py)"""
    root = os.path.join(HERE, "..")
    gen = os.path.join(root, "tools", "gen_eval_check.py")
    for name in ("rv32im", "recursion"):
        irf = os.path.join(HERE, "circuits", name + ".ir.json.gz")
        launcher = os.path.join(CSRC, "gen", "eval_check_%s.cu" % name)
        ptx = glob.glob(os.path.join(CSRC, "gen", "eval_check_%s_p*.ptx" % name))
        newest_in = max([os.path.getmtime(gen), os.path.getmtime(irf),
                         os.path.getmtime(os.path.join(root, "tools", "circuit_ir.py"))] +
                        [os.path.getmtime(f) for f in glob.glob(os.path.join(HERE, "circuits", name + ".tune.json"))])
        if force or not ptx or not os.path.exists(launcher) or min(os.path.getmtime(f) for f in ptx) < newest_in:
            r = subprocess.run([sys.executable, gen, name, "--from-ir"], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("generator failed:\n%s\n%s" % (r.stdout, r.stderr))


def generate_witgen(force=False):
    """(re)create csrc/gen/witgen_rv32im.inc from the committed witgen circuit IR (tools/gen_witgen.py)"""
    root = os.path.join(HERE, "..")
    gen = os.path.join(root, "tools", "gen_witgen.py")
    irf = os.path.join(HERE, "circuits", "rv32im_witgen.ir.json.gz")
    out = os.path.join(CSRC, "gen", "witgen_rv32im.inc")
    # the generator leaves an unchanged output file alone (the step kernels take minutes to rebuild), so "generator ran
    # for these inputs" is recorded in a stamp rather than read off the output's time
    stamp = os.path.join(OBJ, "witgen_gen.stamp")
    key = repr((os.path.getmtime(gen), os.path.getmtime(irf)))
    if force or not os.path.exists(out) or not os.path.exists(stamp) or open(stamp).read() != key:
        r = subprocess.run([sys.executable, gen], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("witgen generator failed:\n%s\n%s" % (r.stdout, r.stderr))
        os.makedirs(OBJ, exist_ok=True)
        with open(stamp, "w") as f:
            f.write(key)


# The two kernels that instantiate the generated rv32im step functions are 19 k lines of branchy straight-line code:
# ptxas -O3 needs 10+ minutes for them and gains nothing measurable (they are latency / divergence bound), -O1 needs two.
SLOW_SOURCES = {"witgen_step_exec.cu": ["-Xptxas", "-O1", "-diag-suppress", "550"],
                "witgen_step_accum.cu": ["-Xptxas", "-O1", "-diag-suppress", "550"]}


def _ptxas(ptx, cubin, flags):
    # ptxas ignores -maxrregcount for entries that carry .maxntid (all generated kernels do): a register cap has to be
    # the kernel's own .maxnreg directive, so the flag is turned into one on a temporary copy of the PTX.
    cap = [f for f in flags if f.startswith("-maxrregcount=")]
    src = ptx
    if cap:
        flags = [f for f in flags if f not in cap]
        text = open(ptx).read()
        if ".maxnreg" not in text:
            text = text.replace("\n{", "\n.maxnreg %d\n{" % int(cap[-1].split("=")[1]), 1)
        src = cubin[:-6] + ".capped.ptx"
        with open(src, "w") as f:
            f.write(text)
    cmd = [os.path.join(os.path.dirname(NVCC), "ptxas"), "-arch=sm_100a", "-v"] + flags + [src, "-o", cubin]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("ptxas failed for %s:\n%s\n%s" % (ptx, r.stdout, r.stderr))
    import re
    m = re.search(r"(\d+) bytes spill stores", r.stderr)
    regs = re.search(r"Used (\d+) registers", r.stderr)
    return int(m.group(1)) if m else 0, int(regs.group(1)) if regs else 0, r.stderr


_TUNED = None


def _tuned_flags():
    """kernel name -> ptxas flags, the fastest of the measured flag sets for that part (risc0_b200/circuits/*.tune.json,
    written by tools/autotune_eval_check.py from per-part device times at po2 = 20). All flag sets assemble the same PTX,
    so results are bit-identical whichever is chosen."""
    global _TUNED
    if _TUNED is None:
        import json
        _TUNED = {}
        if os.environ.get("R0B200_NO_TUNED") != "1":
            for f in glob.glob(os.path.join(HERE, "circuits", "*.tune.json")):
                _TUNED.update(json.load(open(f))["flags"])
    return _TUNED


SPILL_LIMIT = 4000  # bytes of spill stores per thread above which the -O3 schedule loses to the -O1 one (measured)


def _assemble(ptx, cubin, verbose):
    """Generated straight-line kernels: ptxas -O3 schedules them well when the register pressure of the part is low,
    but on the high-pressure parts its pre-allocation scheduler hoists loads across tens of thousands of instructions
    and the allocator falls back to spilling nearly everything. Those parts are assembled with -O1 (program order
    kept). This heuristic only applies to parts without a measured entry in circuits/*.tune.json. Measured per part in gpurun_out/evalcheck_variants*.log."""
    if PTXAS_OPT != "-O3" or PTXAS_EXTRA:          # experiment builds: exactly the flags asked for, for every part
        spill, regs, log = _ptxas(ptx, cubin, [PTXAS_OPT] + PTXAS_EXTRA)
        return ptx, log
    tuned = _tuned_flags().get(os.path.basename(ptx)[:-4])
    if tuned is not None:                          # measured per part on the B200 (tools/autotune_eval_check.py)
        spill, regs, log = _ptxas(ptx, cubin, tuned)
        return ptx, log
    spill, regs, log = _ptxas(ptx, cubin, ["-O3"])
    if spill > SPILL_LIMIT or regs <= 64:
        spill1, regs1, log1 = _ptxas(ptx, cubin, ["-O1"])
        log += "-> re-assembled with -O1:\n" + log1
    return ptx, log


def _compile(src, obj, verbose):
    cmd = [NVCC] + FLAGS + SLOW_SOURCES.get(os.path.basename(src), []) + ["-I", CSRC, "-c", src, "-o", obj]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return src, r.stderr


def build(force=False, verbose=False, jobs=None):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    generate(force)
    generate_witgen(force)
    wg_only = {"witgen_rt.cuh", "witgen_rv32im.inc"}   # headers only the witgen translation units include
    hdr_time = max(os.path.getmtime(h) for h in headers() if os.path.basename(h) not in wg_only)
    wg_time = max(os.path.getmtime(h) for h in headers() + [os.path.join(CSRC, "gen", "witgen_rv32im.inc")])
    todo, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.relpath(src, CSRC).replace(os.sep, "_")[:-3] + ".o")
        objs.append(obj)
        base = os.path.basename(src)
        if base in SLOW_SOURCES:      # include only the witgen runtime, fp.cuh and the generated step functions
            dep_time = max(os.path.getmtime(os.path.join(CSRC, f)) for f in ("witgen_rt.cuh", "fp.cuh", "gen/witgen_rv32im.inc"))
        else:
            dep_time = wg_time if base.startswith("witgen") else hdr_time
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), dep_time):
            todo.append((src, obj))
    todo.sort(key=lambda so: os.path.basename(so[0]) not in SLOW_SOURCES)   # start the slow ones first
    # generated PTX kernels -> cubins, embedded as read-only data (symbol r0_cubin_<file stem>)
    ptx_todo, cubins = [], []
    # the ptxas flag set is part of the cubin staleness key
    stamp = os.path.join(OBJ, "ptxas_flags.stamp")
    flag_key = repr((PTXAS_OPT, PTXAS_EXTRA, SPILL_LIMIT, sorted(_tuned_flags().items())))
    flags_changed = not os.path.exists(stamp) or open(stamp).read() != flag_key
    for ptx in ptx_sources():
        cubin = os.path.join(OBJ, os.path.basename(ptx)[:-4] + ".cubin")
        cubins.append(cubin)
        if force or flags_changed or not os.path.exists(cubin) or os.path.getmtime(cubin) < os.path.getmtime(ptx):
            ptx_todo.append((ptx, cubin))
    if (todo or ptx_todo) and os.path.exists(LIB):
        os.remove(LIB)   # never leave a stale library behind a failed rebuild
    if todo or ptx_todo:
        with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            futs = [ex.submit(_compile, s_, o_, verbose) for s_, o_ in todo]
            futs += [ex.submit(_assemble, p_, c_, verbose) for p_, c_ in ptx_todo]
            for f in futs:
                src, log = f.result()
                if verbose:
                    sys.stderr.write("== %s\n%s" % (os.path.basename(src), log))
    with open(stamp, "w") as f:
        f.write(flag_key)
    if cubins:
        embed_s = os.path.join(OBJ, "embed_cubins.S")
        embed_o = os.path.join(OBJ, "embed_cubins.o")
        text = ['.section .rodata']
        for cubin in cubins:
            sym = "r0_cubin_" + os.path.basename(cubin)[:-6]
            text += [".global %s" % sym, ".balign 16", "%s:" % sym, '.incbin "%s"' % cubin]
        text.append('.section .note.GNU-stack,"",@progbits')
        new = "\n".join(text) + "\n"
        if ptx_todo or not os.path.exists(embed_o) or not os.path.exists(embed_s) or open(embed_s).read() != new:
            with open(embed_s, "w") as f:
                f.write(new)
            r = subprocess.run(["gcc", "-c", embed_s, "-o", embed_o], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("embedding cubins failed:\n%s" % r.stderr)
            todo.append((embed_s, embed_o))
        objs.append(embed_o)
    if todo or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
