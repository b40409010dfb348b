"""Alias so that ``import paillier_b200`` works: the real package lives in the directory
``python-paillier_b200/`` (hyphenated, mirroring the reference repo name)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("python-paillier_b200")
sys.modules[__name__] = _pkg
