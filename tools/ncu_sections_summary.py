#!/usr/bin/env python3
"""Raw `ncu -i rep --page raw --csv` export of a sections capture -> per-launch JSON + markdown table.

    python tools/ncu_sections_summary.py gpurun_out/rN_sections_raw.csv profiles/rN_ncu_sections.json > table.md

Columns: device time, DRAM bytes (dram__bytes.sum.per_second x duration) and %, fma-heavy / alu pipe cycles active,
issue slots, warps active, registers, L1 / L2 hit rates, and the per-issue stall reasons."""
import csv
import json
import re
import sys

COLS = [("ms", "gpu__time_duration.sum", 1e-6), ("dram_Bps", "dram__bytes.sum.per_second", 1.0),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("fmaheavy", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("alu", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("issue", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1.0),
        ("warps", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
        ("regs", "launch__registers_per_thread", 1.0),
        ("l1hit", "l1tex__t_sector_hit_rate.pct", 1.0), ("l2hit", "lts__t_sector_hit_rate.pct", 1.0)]
STALLS = ["long_scoreboard", "math_pipe_throttle", "no_instruction", "wait", "barrier", "not_selected", "short_scoreboard",
          "mio_throttle", "lg_throttle"]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        rec = {"kernel": name}
        for key, col, scale in COLS:
            if col in idx:
                v = num(r[idx[col]])
                u = units[idx[col]]
                if v is not None:
                    if key == "ms":     # duration comes in ns / us / ms depending on the export
                        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
                    elif key == "dram_Bps":
                        v *= {"byte/s": 1.0, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}.get(u, 1.0)
                    rec[key] = v
        for s in STALLS:
            col = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if col in idx:
                rec[s] = num(r[idx[col]])
        if "dram_Bps" in rec and "ms" in rec:
            rec["dram_GB"] = round(rec.pop("dram_Bps") * rec["ms"] * 1e-3 / 1e9, 4)
        out.append(rec)
    json.dump(out, open(sys.argv[2], "w"), indent=0)
    keys = ["ms", "dram_GB", "dram_pct", "fmaheavy", "alu", "issue", "warps", "regs", "l1hit", "l2hit"] + STALLS[:6]
    print("| kernel | " + " | ".join(keys) + " |")
    print("|---|" + "---|" * len(keys))
    for rec in out:
        print("| `%s` | " % rec["kernel"] + " | ".join(("%.3f" % rec[k] if k in ("ms",) else "%.1f" % rec[k] if isinstance(rec.get(k), float) else "-")
                                                   for k in keys) + " |")


if __name__ == "__main__":
    main()
