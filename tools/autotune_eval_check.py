#!/usr/bin/env python3
"""Per-part choice of generator options and ptxas flags for the generated eval_check kernels, from measured times.

    tools/autotune_build.sh                       # here: every part under every variant of tools/autotune_variants.json
    gpurun -- 'bash tools/autotune_run.sh'        # B200: per-part device times of each variant -> gpurun_out/r2_autotune.log
    python tools/autotune_eval_check.py gpurun_out/r2_autotune.log   # -> risc0_b200/circuits/rv32im.tune.json

Every variant computes the same exact field arithmetic (only instruction order, re-load distance, fence distance,
register cap and ptxas level differ), so results are bit-identical whichever variant a part gets; the log carries the
CRC of `check` per run as a cross-check. tools/gen_eval_check.py reads the "gen" table, risc0_b200/build.py "flags".
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    log = sys.argv[1]
    variants = json.load(open(os.path.join(ROOT, "risc0_b200", "lib", "cubins_at", "variants.json")))
    runs = [json.loads(l) for l in open(log) if l.startswith("{")]
    base = [r for r in runs if r["lib"] == "default"][0]
    crcs = {r["check_crc32"] for r in runs}
    assert len(crcs) == 1, "variants disagree on the result: %s" % crcs
    by_var = {r["lib"]: r for r in runs if r["lib"] in variants}
    parts = sorted((k for k in base if k.startswith("eval_check_p")), key=lambda k: int(k[12:]))
    gen, flags, total, per_part = {}, {}, 0.0, {}
    for p in parts:
        best = min(by_var, key=lambda v: by_var[v][p])
        t_best, t_base = by_var[best][p], base[p]
        name = "eval_check_rv32im_" + p[11:]
        # keep the default unless the gain is real (> 1.5 %): run-to-run noise is ~0.5 %
        if t_best < 0.985 * t_base:
            if variants[best]["gen"]:
                gen[name] = variants[best]["gen"]
            flags[name] = variants[best]["flags"]
            total += t_best
            chosen = best
        else:
            total += t_base
            chosen = "default"
        per_part[p] = dict(default_ms=t_base, best_ms=t_best, best=best, chosen=chosen,
                           top5={v: by_var[v][p] for v in sorted(by_var, key=lambda v: by_var[v][p])[:5]})
        print("%-16s default %.3f  best %.3f  %-6s gen %s flags %s" % (p, t_base, t_best, chosen, variants[best]["gen"], " ".join(variants[best]["flags"])))
    if "--update" in sys.argv:      # a later round: `default` in this log is the build tuned so far; keep its entries
        path = os.path.join(ROOT, "risc0_b200", "circuits", "rv32im.tune.json")
        old = json.load(open(path))
        for name in list(old["flags"]):
            if name not in flags:
                flags[name] = old["flags"][name]
                if name in old["gen"]:
                    gen[name] = old["gen"][name]
        for p_, e in old["per_part"].items():
            if per_part[p_]["chosen"] == "default":
                per_part[p_] = dict(e, tuned_ms=per_part[p_]["default_ms"])
            else:
                per_part[p_]["default_ms"] = e["default_ms"]
        base = dict(base, eval_check=old["baseline_ms"])
    out = dict(circuit="rv32im", measured="po2 = %d on a B200, per-part CUDA-event times (tools/autotune_run.sh)" % base["po2"],
               baseline_ms=base["eval_check"], predicted_ms=round(total, 3), gen=gen, flags=flags, per_part=per_part)
    with open(os.path.join(ROOT, "risc0_b200", "circuits", "rv32im.tune.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("baseline %.2f ms -> predicted %.2f ms, %d parts changed" % (base["eval_check"], total, len(flags)))


if __name__ == "__main__":
    main()
