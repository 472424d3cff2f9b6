#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel body (share of total device time)."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = set(sys.argv[2:])
agg = collections.OrderedDict()
for r in rows[1:]:
    m = re.search(r"kernel_entry(?:_lb)?<(?:\d+, \d+, )?(?:dr::)?(\w+)", r[ki])
    short = m.group(1) if m else r[ki].split("(")[0]
    if short in skip:
        continue
    t = float(r[vi].replace(",", ""))
    t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(rows) - 1} launches, {tot:.3f} ms (excluding {sorted(skip)})")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:30s} n={n:3d} {t:10.3f} ms {100 * t / tot:6.2f}%")
