#!/usr/bin/env python3
"""Entry point kept at the repository root for command-line compatibility with the reference's `deploy.py`."""
from aggregathor_b200.cli.deploy import main

if __name__ == "__main__":
  main()
