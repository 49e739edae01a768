#!/usr/bin/env python3
"""Entry point kept at the repository root for command-line compatibility with the reference's `runner.py`."""
import sys

from aggregathor_b200.cli.runner import main

if __name__ == "__main__":
  sys.exit(main())
