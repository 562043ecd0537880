"""Registers the hyphenated package directory under the importable alias ``satk``."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


def load():
    if "satk" not in sys.modules:
        sys.modules["satk"] = importlib.import_module("self-attention-tacotron_b200")
    return sys.modules["satk"]
