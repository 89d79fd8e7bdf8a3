#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import sys


def summarise(fn, out=sys.stdout):
    hdr, agg = None, collections.defaultdict(list)
    for r in csv.reader(open(fn, errors="replace")):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1000 if u.startswith("ns") else (v * 1000 if u.startswith("ms") else v)
        agg[d["Kernel Name"].split("(")[0]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# launches={sum(len(v) for v in agg.values())} total={tot:.1f} us", file=out)
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:44s} n={len(v):5d} total={sum(v):10.1f}us avg={sum(v) / len(v):8.2f}us share={100 * sum(v) / tot:5.1f}%", file=out)


if __name__ == "__main__":
    for f in sys.argv[1:]:
        summarise(f)
