#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time,
share of the captured region.  Usage: python profiles/summarize_launches.py gpurun_out/launches.csv [top_n]"""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=25):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    reader = csv.DictReader(lines)
    for r in reader:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            val = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], val * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for name, us in rows:
        short = re.sub(r"\(.*", "", name)[:110]
        agg[short][0] += 1
        agg[short][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {total / 1e3:.2f} ms of kernel time (cold-cache, serialised: compare shares)")
    print(f"{'kernel':110s} {'n':>6s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{name:110s} {n:6d} {us:12.1f} {us / n:10.2f} {100 * us / total:6.2f}%")
    ours = {k: v for k, v in agg.items() if "b200dp" in k}
    if ours:
        us = sum(v[1] for v in ours.values())
        print(f"\n# b200dp kernels: {sum(v[0] for v in ours.values())} launches, {us:.1f} us, {100 * us / total:.2f}% of kernel time")
        for name, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            print(f"{name:110s} {n:6d} {t:12.1f} {t / n:10.2f} {100 * t / total:6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
