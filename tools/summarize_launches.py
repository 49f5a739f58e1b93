#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (markdown).

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("r0::", "")
        tot[name] += float(r[vi].replace(",", ""))
        cnt[name] += 1
    total = sum(tot.values())
    print("| kernel | launches | total ms | share |")
    print("|---|---:|---:|---:|")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print("| `%s` | %d | %.3f | %.1f %% |" % (k, cnt[k], tot[k] / 1e6, 100 * tot[k] / total))
    print("\ntotal device time in listed launches: %.2f ms over %d launches" % (total / 1e6, sum(cnt.values())))


if __name__ == "__main__":
    main()
