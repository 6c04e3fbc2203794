"""Puts the host package `fddm_b200` (one directory up) on sys.path for the shim modules."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
