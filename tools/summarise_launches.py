#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`):
launches, total and mean duration, share of the repo's own kernel time. Usage: python tools/summarise_launches.py X.csv [header text] > summary.csv
(the files profiles/launches_*_summary.csv were produced this way)."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else path
    rows = list(csv.reader(ln for ln in open(path) if not ln.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, other = collections.OrderedDict(), 0.0
    for r in rows[1:]:
        name, v = r[ki], float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)      # -> microseconds
        if "k_" not in name or "at::" in name or "cutlass" in name:                 # torch / library kernels (problem generation, residual check)
            other += v
            continue
        short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {note}")
    print("kernel,launches,total_us,share_of_engine_time_pct,us_per_launch")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"\"{k}\",{a[0]},{a[1]:.1f},{100 * a[1] / tot:.2f},{a[1] / a[0]:.1f}")
    print(f"# engine total {tot:.1f} us; non-engine (torch RNG problem generation etc.) {other:.1f} us")


if __name__ == "__main__":
    main()
