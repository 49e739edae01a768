#!/usr/bin/env python3
"""Hot spots of a kernel from an .ncu-rep captured with `--set full --import-source on`: the instructions (SASS, with their CUDA source
line when -lineinfo was used) that collect the most warp-stall samples. Usage: ncu_source_top.py file.ncu-rep [rows]"""
import csv
import io
import subprocess
import sys

rows_wanted = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
blocks, current = [], []
for line in out.splitlines():
  if line.startswith('"') or (current and line.strip()):
    current.append(line)
  elif current:
    blocks.append(current)
    current = []
if current:
  blocks.append(current)
for block in blocks:
  reader = list(csv.reader(io.StringIO("\n".join(block))))
  if len(reader) < 3:
    continue
  title = ""
  while reader and "Source" not in reader[0] and "SASS" not in reader[0]:   # leading "Kernel Name", ... rows
    title = reader[0][1] if len(reader[0]) > 1 else title
    reader = reader[1:]
  if len(reader) < 2:
    continue
  header = reader[0]
  if title:
    print("#", title[:160])
  sample_cols = [i for i, h in enumerate(header) if "Sampling" in h and "All" in h] or [i for i, h in enumerate(header) if "Sampl" in h]
  if not sample_cols:
    print("columns:", header[:12])
    continue
  key = sample_cols[0]
  def num(v):
    try:
      return float(v.replace(",", ""))
    except ValueError:
      return 0.0
  body = [r for r in reader[1:] if len(r) == len(header)]
  total = sum(num(r[key]) for r in body) or 1.0
  src = next((i for i, h in enumerate(header) if h.strip() in ("Source", "SASS")), 1)
  stall_cols = [i for i, h in enumerate(header) if h.startswith("stall_") or "Stall" in h]
  print("== %d instructions, %d samples; top %d by %s" % (len(body), int(total), rows_wanted, header[key]))
  for r in sorted(body, key=lambda r: -num(r[key]))[:rows_wanted]:
    reasons = sorted(((num(r[i]), header[i]) for i in stall_cols if num(r[i]) > 0), reverse=True)[:3]
    print("%5.1f%%  %-70s  %s" % (100.0 * num(r[key]) / total, r[src][:70], ", ".join("%s %d" % (h.replace("stall_", ""), v) for v, h in reasons)))
