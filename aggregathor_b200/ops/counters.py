"""Bookkeeping of native kernel launches (the `gpu_launches` figure reported by bench.py)."""

launches = 0


def bump(count=1):
  global launches
  launches += count
