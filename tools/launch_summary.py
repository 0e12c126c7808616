#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total and mean duration, share.
usage: tools/launch_summary.py <launches.csv> [title]   -> markdown on stdout"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    print("### %s\n\n%d launches, %.3f ms of kernel time under the profiler (cold caches, serialised)\n" % (title, sum(a[0] for a in agg.values()), total))
    print("| kernel | launches | total ms | mean ms | share |\n|---|---:|---:|---:|---:|")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.4f | %.1f %% |" % (name, n, ms, ms / n, 100 * ms / total))
    print()


if __name__ == "__main__":
    main()
