"""Device discovery and allocation (reference: `cluster.py:46-221`).

The reference asks a TF server for `/job:J/replica:0/task:T/device:TYPE:id` names and hands devices out to the
workers, the PS and the evaluator, preferring device types in a given order, spreading consecutive allocations over
as many tasks as possible, and letting "reusable" device types host several entities (`--reuse-gpu`).

Here a *task* is one rank of the SPMD job (one process per GPU) and its devices are `GPU:<local index>` (when CUDA is
usable) and `CPU:0`. The job names of the `--server` cluster specification are mapped onto the ranks: the `workers`
job gets one task per rank, `ps` and `eval` live on rank 0 (the PS *role* is executed symmetrically by every rank,
rank 0 is merely the one that reports, evaluates and checkpoints). The allocation algorithm itself — preference
order, task spread, reuse, partial allocation, the report — is the reference's.
"""

from . import tools


class Manager:
  """Inventory of (job, task, device type, device id) and allocation bookkeeping."""

  def __init__(self, structure, devs=None, reuse=None):
    """`structure`: {job: {task id: {device type: [device id, ...]}}}; `devs`: preference order of device types;
    `reuse`: device types that may be allocated several times."""
    self._structure = {}
    self._devs = tuple(devs) if devs is not None else ("GPU", "CPU")
    self._reuse = set(reuse) if reuse is not None else set()
    self._tasks = []  # list of per-task lists of entries [used list, (job, task, type, id)]
    for job, tasks in structure.items():
      self._structure[job] = {}
      for taskid, devtypes in tasks.items():
        self._structure[job][taskid] = {}
        available = []
        for devtype, ids in devtypes.items():
          entries = [{"id": str(devid), "used": []} for devid in ids]
          self._structure[job][taskid][devtype] = entries
          if devtype in self._devs:
            for entry in entries:
              available.append((entry["used"], (job, str(taskid), devtype, entry["id"])))
        if available:
          self._tasks.append(available)

  @classmethod
  def from_world(cls, world, gpus_per_rank, spec=None, ps_job="ps", wk_job="workers", ev_job="eval", devs=None, reuse=None):
    """Inventory of an SPMD job of `world` ranks. `gpus_per_rank[r]` = CUDA device index of rank r, or None (CPU only)."""
    def devices(rank):
      out = {"CPU": ["0"]}
      if gpus_per_rank[rank] is not None:
        out["GPU"] = [str(gpus_per_rank[rank])]
      return out
    structure = {}
    for job in dict.fromkeys((wk_job, ps_job, ev_job)):
      if job == wk_job:
        structure[job] = {rank: devices(rank) for rank in range(world)}
      else:
        structure.setdefault(job, {0: devices(0)})
    if spec is not None:
      for job in spec:
        structure.setdefault(job, {0: devices(0)})
    return cls(structure, devs, reuse)

  def report(self):
    print("Cluster structure and allocation report:")
    for job, tasks in self._structure.items():
      print(" · Job " + repr(job))
      for taskid, devtypes in tasks.items():
        print("   · Task " + repr(taskid))
        for devtype, entries in devtypes.items():
          for entry in entries:
            print("     · " + devtype + " " + entry["id"] + ": " + ("<unallocated>" if not entry["used"] else ", ".join(entry["used"])))

  def allocate(self, name, count, jobs=None, partial=False):
    """Reserve `count` devices for `name`; returns [(job, task, type, id)] or None when impossible (unless `partial`)."""
    if count == 0:
      tools.warning("Successfully allocated 0 device")
      return []
    if count < 0:
      raise tools.UserException("Expected non-negative number of devices to reserve, got " + repr(count))
    if jobs is not None:
      for job in jobs:
        if job not in self._structure:
          tools.warning("Job " + repr(job) + " does not exist in the cluster")
    tasks = [list(task) for task in self._tasks]
    chosen = []
    for devtype in self._devs:
      progress = True
      while progress and len(chosen) < count:
        progress = False
        for position, task in enumerate(tasks):
          pick = None
          for index, (used, info) in enumerate(task):
            if info[2] != devtype or (jobs is not None and info[0] not in jobs):
              continue
            if devtype not in self._reuse and (used or any(c[1] == info for c in chosen)):
              continue
            pick = index
            break
          if pick is None:
            continue
          entry = task.pop(pick)
          if devtype in self._reuse:
            task.append(entry)  # stays available, but behind the task's other devices
          tasks.append(tasks.pop(position))  # a task that just gave a device goes to the back: maximise spread
          chosen.append(entry)
          progress = True
          break
      if len(chosen) == count:
        break
    if len(chosen) < count and not partial:
      return None
    self._tasks = [task for task in tasks if task]
    result = []
    for index, (used, info) in enumerate(chosen):
      used.append(name + "[" + str(index) + "]")
      result.append(info)
    return result
