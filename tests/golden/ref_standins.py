"""Kept for `make_golden.py`: the stand-ins that let the unmodified reference import moved to baseline/ref_harness.py
(bench.py's reference arms use them too)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from baseline.ref_harness import *  # noqa: F401,F403,E402
from baseline.ref_harness import install, make_cfg, REFERENCE_ROOT  # noqa: F401,E402
