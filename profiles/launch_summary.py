#!/usr/bin/env python3
"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:
   python profiles/launch_summary.py profiles/r01b_launches_bench.csv > profiles/r01b_launches_summary.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py, not absolutes."""
import collections
import csv
import sys

SCALE = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].split("::")[-1]
        ms = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:32s} launches={n:4d} total={t:10.3f} ms  avg={t / n:9.4f} ms  share={100 * t / tot:5.1f}%")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
