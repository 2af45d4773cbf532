#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total device time,
launches and share of the step per kernel.  usage: tools/summarize_launches.py launches.csv"""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, agg = None, {}
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(d["Metric Value"].replace(",", ""))
            v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(d["Metric Unit"], 1.0)
            name = d["Kernel Name"]
            name = name[:name.index("(")] if "(" in name else name
            agg.setdefault(name[:90], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, {tot / 1e6:.3f} ms of device time "
          "(cold-cache, serialised under ncu: compare shares, not absolutes)")
    print(f"{'total ms':>10} {'launches':>8} {'share':>7}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v) / 1e6:10.3f} {len(v):8d} {100 * sum(v) / tot:6.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
