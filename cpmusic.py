"""Import alias: ``import cpmusic`` -> the package in
``reinforcement-learning-in-music-generation_b200/`` (a directory name Python's import
statement cannot spell)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("reinforcement-learning-in-music-generation_b200")
sys.modules[__name__] = _pkg
