#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares of device time)."""
import csv, sys, re, collections

def main(path, top=40):
    lines = [l for l in open(path, errors="ignore") if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        unit = r[ui]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r[ki])[:90]
        agg[name][0] += 1
        agg[name][1] += ns
    total = sum(v[1] for v in agg.values())
    print(f"# kernel launch list `{path}`: {sum(v[0] for v in agg.values())} launches, {total/1e6:.2f} ms total (serialised, cold cache: compare shares)\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| `{name}` | {n} | {ns/1e6:.3f} | {100*ns/total:.1f}% |")

if __name__ == "__main__":
    main(sys.argv[1])
