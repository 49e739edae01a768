#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into a few roofline-relevant lines per kernel launch."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum", "smsp__cycles_active.avg",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__inst_executed.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
header = rows[0]
units = rows[1]
for row in rows[2:]:
  rec = dict(zip(header, row))
  name = rec.get("Kernel Name", "?")[:100]
  print("=" * 20, rec.get("ID"), name)
  for key in KEYS:
    if key in rec:
      print("   %-72s %s %s" % (key, rec[key], units[header.index(key)]))
