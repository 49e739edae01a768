"""Measurement utilities: `clocks` (nvidia-smi sampling around timed regions), `timing` (CUDA-event timing, max over ranks).
The logging / registry / checkpoint helpers that mirror the reference's `tools` package live in `aggregathor_b200.tools`."""
