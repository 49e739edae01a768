#!/usr/bin/env python3
"""Per-kernel SASS listings of the native libraries (`cuobjdump -sass`, encodings stripped) -> profiles/sass/<library>/<kernel>.sass,
plus profiles/sass/INDEX.txt with every kernel's instruction count and its Blackwell-specific mnemonics (UTC*MMA = tcgen05.mma,
LDTM = tcgen05.ld, UTMALDG = TMA tensor loads, UTCBAR = tcgen05.commit, SYNCS = mbarrier, LDGMC/STGMC... = multimem, *.SYS = system-scope
peer accesses). Runs on the CPU box (no GPU needed). Usage: sass_listing.py [--all] (default: the kernels named in KEY)."""

import pathlib
import re
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
NATIVE = ROOT / "aggregathor_b200" / "native"
OUT = ROOT / "profiles" / "sass"
KEY = [r"gemm_tcgen05_persistent_kernel<128, 0, 0, ElemBF16>", r"gemm_tcgen05_persistent_kernel<128, 1, 1, ElemBF16>", r"gemm_tcgen05_persistent_kernel<128, 0, 0, ElemTF32>",
       r"gemm_tcgen05_persistent_kernel<128, 1, 1, ElemTF32>", r"gemm_tcgen05_pair_kernel<256>", r"conv_tcgen05_kernel<128, 0, ElemBF16>", r"conv_tcgen05_kernel<128, 1, ElemBF16>",
       r"conv_tcgen05_kernel<128, 2, ElemBF16>", r"conv_tcgen05_kernel<128, 0, ElemTF32>", r"gar_fused_kernel<8, 4>", r"gar_phase_a_kernel<0>", r"bn_fused_kernel<0>", r"bn_fused_kernel<1>",
       r"preprocess_kernel<__nv_bfloat16>", r"allreduce_kernel<float, 1>", r"allgather_kernel", r"channel_sums_kernel<float, 0, 32>", r"depthwise_fwd_kernel"]
SPECIAL = re.compile(r"\b(UTC[A-Z]*MMA[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|SYNCS[.\w]*|LDGMC[.\w]*|STGMC[.\w]*|REDGMC[.\w]*|UCGABAR[.\w]*|MAPA[.\w]*|"
                     r"LDG\.E[.\w]*SYS|STG\.E[.\w]*SYS|LD\.E[.\w]*SYS|ST\.E[.\w]*SYS|RED\.E[.\w]*|HMMA[.\w]*)")


def demangle(names):
  proc = subprocess.run(["cu++filt"] + names, stdout=subprocess.PIPE, text=True)
  return proc.stdout.strip().split("\n") if proc.returncode == 0 else names


def main():
  write_all = "--all" in sys.argv
  OUT.mkdir(parents=True, exist_ok=True)
  index = []
  for so in sorted(NATIVE.glob("op_*.so")):
    text = subprocess.run(["cuobjdump", "-sass", str(so)], stdout=subprocess.PIPE, text=True).stdout
    chunks = re.split(r"\n\s*Function : ", text)[1:]
    mangled = [chunk.split("\n", 1)[0].strip() for chunk in chunks]
    names = demangle(mangled)
    for name, chunk in zip(names, chunks):
      lines = []
      for line in chunk.split("\n")[1:]:
        line = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip()
        if line.strip() and not re.fullmatch(r"\s*", line):
          lines.append(line)
      body = [l for l in lines if re.search(r"/\*[0-9a-f]{4}\*/", l)]
      counts = {}
      for l in body:
        for m in SPECIAL.findall(l):
          counts[m] = counts.get(m, 0) + 1
      short = re.sub(r"^void |\(anonymous namespace\)::|<unnamed>::", "", name)
      short = re.sub(r"\((int|bool|unsigned int)\)", "", short)
      short = short.split("(")[0]
      index.append((so.stem, short, len(body), counts))
      if write_all or any(re.search(pattern, short) for pattern in KEY):
        target = OUT / so.stem / (re.sub(r"[^A-Za-z0-9_.<>,-]+", "_", short)[:150] + ".sass")
        target.parent.mkdir(parents=True, exist_ok=True)
        target.write_text("// " + name + "\n// " + so.name + ", sm_100a, cuobjdump -sass (instruction encodings stripped)\n" + "\n".join(lines) + "\n")
  with open(OUT / "INDEX.txt", "w") as fd:
    fd.write("library | kernel | SASS instructions | Blackwell / cross-GPU mnemonics (count)\n")
    for lib, short, count, counts in index:
      fd.write("%s | %s | %d | %s\n" % (lib, short, count, ", ".join("%s x%d" % kv for kv in sorted(counts.items()))))
  print("indexed", len(index), "kernels")


if __name__ == "__main__":
  main()
