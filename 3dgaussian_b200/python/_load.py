"""Locates and imports the `3dgaussian_b200` package for the drop-in shim modules."""
import importlib
import os
import sys

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOT = os.path.dirname(_PKG_DIR)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
pkg = importlib.import_module(os.path.basename(_PKG_DIR))
renderer = importlib.import_module(os.path.basename(_PKG_DIR) + ".renderer")
