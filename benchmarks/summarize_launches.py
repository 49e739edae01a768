#!/usr/bin/env python3
"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum` log by kernel name: calls, total time, share."""
import collections
import csv
import re
import sys

rows = collections.OrderedDict()
total = 0.0
with open(sys.argv[1], newline="") as fd:
  lines = [l for l in fd if l.startswith('"')]
reader = csv.DictReader(lines)
for row in reader:
  if row.get("Metric Name") != "gpu__time_duration.sum":
    continue
  name = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
  value = float(row["Metric Value"].replace(",", ""))
  unit = row.get("Metric Unit", "ns")
  ns = value * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
  entry = rows.setdefault(name, [0, 0.0])
  entry[0] += 1
  entry[1] += ns
  total += ns
print("total device time %.3f ms over %d launches" % (total / 1e6, sum(v[0] for v in rows.values())))
for name, (count, ns) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
  print("%6.2f%% %9.3f ms %5d  %s" % (100.0 * ns / total, ns / 1e6, count, name))
