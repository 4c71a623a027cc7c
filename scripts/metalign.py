#! /usr/bin/env python
"""Same name and command line as scripts/metalign.py of the reference; runs metalign_b200.metalign."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metalign_b200.metalign import main  # noqa: E402

if __name__ == "__main__":
    main()
