#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, mean time and share."""
import csv, re, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("isph::", "")
    v = float(r[iv].replace(",", "")); v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(f"# launches {sum(cnt.values())}  total {T / 1e3:.1f} ms")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{100 * tot[k] / T:6.2f}%  {cnt[k]:6d} launches  avg {tot[k] / cnt[k]:9.1f} us  {k}")
