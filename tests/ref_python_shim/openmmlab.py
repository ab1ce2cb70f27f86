"""STAND-IN for the reference's SWIG module `openmmlab`: the host mirror's class under the reference's module name."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from openmmlab_b200 import SlicedNonbondedForce  # noqa: E402,F401
