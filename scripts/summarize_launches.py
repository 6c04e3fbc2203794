#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    k = row["Kernel Name"][:78]
    tot[k] += float(row["Metric Value"].replace(",", ""))
    cnt[k] += 1
T = sum(tot.values())
print(f"{sum(cnt.values())} launches, {T / 1e3:.1f} us total")
for k, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    print(f"{v / 1e3:10.1f} us {cnt[k]:4d}x {v / cnt[k] / 1e3:9.2f} us/launch {100 * v / T:5.1f}%  {k}")
