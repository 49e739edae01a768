"""aggregathor_b200 — Byzantine-resilient synchronous data-parallel training for one box of B200 GPUs, with the capabilities and the
command-line surface of LPD-EPFL/AggregaThor (see README.md, DESIGN.md, COVERAGE.md)."""

__version__ = "0.1.0"
