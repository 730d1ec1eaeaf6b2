#!/usr/bin/env python
"""Per-kernel totals and shares from an ncu launch list (--metrics gpu__time_duration.sum --csv). Optionally only the
last K steps' worth of launches (--tail N launches)."""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    tail = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = rows[1:]
    if tail:
        data = data[-tail:]
    agg = OrderedDict()
    for r in data:
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        t = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += t
    tot = sum(a[1] for a in agg.values())
    print("%d launches, %.3f ms total (ncu per-launch times: cold-cache, serialised -- compare shares)" % (len(data), tot / 1e6))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s n=%4d  total=%10.3f ms  mean=%9.3f ms  share=%5.1f%%" % (name[:60], n, t / 1e6, t / 1e6 / n, 100 * t / tot))


if __name__ == "__main__":
    main()
