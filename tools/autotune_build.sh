#!/bin/bash
# Generates the eval_check PTX under every generator option set of tools/autotune_variants.json and assembles every
# part with every ptxas flag set into risc0_b200/lib/cubins_at/<variant>/ (loaded through R0B200_CUBIN_DIR by the
# launcher). The partition (which terms a part holds) is the same in all of them.
set -e
cd "$(dirname "$0")/.."
rm -rf risc0_b200/lib/cubins_at /tmp/at_gen
python - <<'PY'
import concurrent.futures as cf, glob, json, os, subprocess, sys
sys.path.insert(0, ".")
from risc0_b200 import build as b
V = json.load(open(os.environ.get("AT_VARIANTS", "tools/autotune_variants.json")))
PFX = V.get("prefix", "g")
variants = []   # (name, gen env, flags)
for gi, g in enumerate(V["gen"]):
    for fi, f in enumerate(V["flags"]):
        variants.append(("%s%df%d" % (PFX, gi, fi), g, f))
for fi, f in enumerate(V["extra_default_gen_flags"]):
    variants.append(("%s0x%d" % (PFX, fi), {}, f))
gens = {}
for gi, g in enumerate(V["gen"]):
    d = "/tmp/at_gen/g%d" % gi
    os.makedirs(d, exist_ok=True)
    env = dict(os.environ, EVAL_IGNORE_TUNE="1", EVAL_OUT_DIR=d, **g)
    gens[gi] = (d, subprocess.Popen([sys.executable, "tools/gen_eval_check.py", "rv32im", "--from-ir"], env=env, stdout=subprocess.DEVNULL))
for d, p in gens.values():
    assert p.wait() == 0
    assert len(glob.glob(d + "/*.ptx")) == 36, d
jobs = []
with cf.ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
    for name, g, f in variants:
        gi = V["gen"].index(g) if g in V["gen"] else None
        if gi is None:      # default generator options (extra flag sets): generate once
            gi = "d"
            if gi not in gens:
                d = "/tmp/at_gen/gd"
                os.makedirs(d, exist_ok=True)
                subprocess.check_call([sys.executable, "tools/gen_eval_check.py", "rv32im", "--from-ir"], env=dict(os.environ, EVAL_IGNORE_TUNE="1", EVAL_OUT_DIR=d), stdout=subprocess.DEVNULL)
                gens[gi] = (d, None)
        out = os.path.join("risc0_b200", "lib", "cubins_at", name)
        os.makedirs(out, exist_ok=True)
        for ptx in sorted(glob.glob(gens[gi][0] + "/eval_check_rv32im_p*.ptx")):
            jobs.append(ex.submit(b._ptxas, ptx, os.path.join(out, os.path.basename(ptx)[:-4] + ".cubin"), list(f)))
    for j in jobs:
        j.result()
for f in glob.glob("risc0_b200/lib/cubins_at/*/*.capped.ptx"):
    os.remove(f)
json.dump({n: dict(gen=g, flags=f) for n, g, f in variants}, open("risc0_b200/lib/cubins_at/variants.json", "w"), indent=1)
print(len(variants), "variants,", len(jobs), "cubins")
PY
