"""Import shim: ``import dsvt_b200`` == the package in ``dsvt-ai-trt_b200/`` (hyphenated name)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("dsvt-ai-trt_b200")
sys.modules[__name__] = _pkg
