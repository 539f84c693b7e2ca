"""Importable alias: `import mic_b200` == the package directory `multilingual-image-captioning_b200/`
(a hyphenated directory name cannot be written in an `import` statement)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("multilingual-image-captioning_b200")
sys.modules[__name__] = _pkg
