#!/usr/bin/env python
"""Drop-in for the reference entry point: `python src/main.py --env-config=<env> --config=<alg> [with k=v ...]`
(reference: src/main.py:66-102).  The implementation lives in refil_b200/main.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from refil_b200.main import main  # noqa: E402

if __name__ == "__main__":
    main()
