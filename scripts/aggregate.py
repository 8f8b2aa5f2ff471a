#!/usr/bin/env python
"""CLI kept at the reference's path: ``python scripts/aggregate.py -b DIR -o OUT [-d]``."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megatronapp_b200.scan.aggregate import main  # noqa: E402

if __name__ == "__main__":
    main()
