"""GPU clock / throttle-reason sampling around a timed region (the clocks line of the B200 profiling recipe)."""

import statistics
import subprocess


class ClockSampler:
  """`nvidia-smi` clocks/throttle-reason sampling during the timed region (profiling recipe's clocks line)."""

  QUERY = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

  def __init__(self, index):
    self.index, self.proc = index, None

  def start(self):
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
      self.proc = None

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      out, _ = self.proc.communicate(timeout=5)
    except Exception:
      self.proc.kill()
      out = ""
    clocks, maxes, reasons = [], [], set()
    for line in out.splitlines():
      fields = [f.strip() for f in line.split(",")]
      if len(fields) < 9:
        continue
      try:
        clocks.append(float(fields[1]))
        maxes.append(float(fields[2]))
      except ValueError:
        continue
      for name, value in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), fields[5:9]):
        if value.lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": statistics.median(clocks) if clocks else None, "sm_max_mhz": max(maxes) if maxes else None, "reasons": sorted(reasons), "samples": len(clocks)}
