#! /usr/bin/env python
"""GPU sketch builder: genome FASTA files -> .mlgdb (see metalign_b200/sketch.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metalign_b200.sketch import main  # noqa: E402

if __name__ == "__main__":
    main()
