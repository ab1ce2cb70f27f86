"""Importable alias: the package directory is named ``openmm-lab_b200`` (not a valid
Python identifier), so ``import openmmlab_b200`` loads it through importlib and aliases
the package and every submodule (one module object per file, whichever name is used)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_real = "openmm-lab_b200"
_pkg = importlib.import_module(_real)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_real+"."):
        sys.modules[__name__+"."+_name[len(_real)+1:]] = _mod
sys.modules[__name__] = _pkg
