"""Thin alias of ``oracle/ref_shim.py`` (kept so ``make_golden.py`` and older scripts keep importing it).
TEST INFRASTRUCTURE ONLY."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle.ref_shim import import_reference, reference_root  # noqa: E402,F401
