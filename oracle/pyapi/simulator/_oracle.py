"""Locates ``oracle/binding.py`` from inside the stand-in package."""
import os
import sys

_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import binding  # noqa: E402

__all__ = ["binding"]
