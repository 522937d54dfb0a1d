"""Importable alias for the package directory (whose name contains hyphens)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("multimodal-framework-for-speaker-emotion-recognition_b200")
sys.modules[__name__] = _pkg
