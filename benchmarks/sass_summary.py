#!/usr/bin/env python3
"""Per-kernel SASS evidence: counts of the mnemonics that prove the Blackwell-native paths (tcgen05 -> UTC*MMA / LDTM,
TMA -> UTMALDG, mbarrier -> SYNCS, NVLS -> MULTIMEM..., system-scope flags -> *.SYS), plus registers / shared memory.
Runs on the CPU box: python benchmarks/sass_summary.py > profiles/sass_summary.txt"""
import collections
import pathlib
import re
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent.parent
NATIVE = ROOT / "aggregathor_b200" / "native"
INTERESTING = re.compile(r"\b(UTC[A-Z0-9]*MMA[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|UTCATOMSWS[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|STTM[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UTMASTG[A-Z0-9_.]*|UTMAPF[A-Z0-9_.]*|UBLKCP[A-Z0-9_.]*|SYNCS[A-Z0-9_.]*|MULTIMEM[A-Z0-9_.]*|LDGMC[A-Z0-9_.]*|STGMC[A-Z0-9_.]*|REDGMC[A-Z0-9_.]*|STMC[A-Z0-9_.]*|"
                         r"LDG\.E[A-Z0-9_.]*SYS[A-Z0-9_.]*|STG\.E[A-Z0-9_.]*SYS[A-Z0-9_.]*|LD\.E[A-Z0-9_.]*SYS[A-Z0-9_.]*|ST\.E[A-Z0-9_.]*SYS[A-Z0-9_.]*|MEMBAR[A-Z0-9_.]*SYS[A-Z0-9_.]*|"
                         r"RED\.E[A-Z0-9_.]*|ATOMG[A-Z0-9_.]*|HMMA[A-Z0-9_.]*|LDGSTS[A-Z0-9_.]*|FENCE[A-Z0-9_.]*)")

for lib in sorted(NATIVE.glob("op_*.so")):
  sass = subprocess.run(["cuobjdump", "-sass", str(lib)], stdout=subprocess.PIPE, text=True).stdout
  res = subprocess.run(["cuobjdump", "-res-usage", str(lib)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
  usage = {}
  current = None
  for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
      current = m.group(1)
      continue
    if current and "REG:" in line:
      usage[current] = line.strip()
      current = None
  print("=" * 100)
  print(lib.name)
  kernel, counts = None, {}
  order = []
  for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
      kernel = m.group(1)
      counts[kernel] = collections.Counter()
      order.append(kernel)
      continue
    if kernel is None:
      continue
    for hit in INTERESTING.findall(line):
      counts[kernel][hit] += 1
  for kernel in order:
    demangled = subprocess.run(["c++filt", kernel], stdout=subprocess.PIPE, text=True).stdout.strip()
    demangled = re.sub(r"\(anonymous namespace\)::", "", demangled)
    demangled = re.sub(r"\(.*", "", demangled)[:110]
    print("-" * 100)
    print(demangled)
    if kernel in usage:
      print("   " + usage[kernel])
    if counts[kernel]:
      print("   " + ", ".join("%s x%d" % (k, v) for k, v in sorted(counts[kernel].items())))
