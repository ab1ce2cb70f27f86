"""STAND-IN for the `openmm` Python package, just wide enough for the reference's OWN Python tests of SlicedNonbondedForce
(python/tests/TestSlicedNonbondedForce.py) to run UNCHANGED on this repository's host mirror (tests/test_reference_python_tests.py).
Test infrastructure: platform "Reference" is the CPU oracle, "CUDA" the B200 kernel; a plain NonbondedForce is evaluated as a
one-subset SlicedNonbondedForce (OpenMM's own nonbonded kernels are not in this image)."""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import openmmlab_b200 as _mm  # noqa: E402
from oracle import oracle_kernel as _oracle  # noqa: E402,F401  (registers the CPU platforms)

from . import unit  # noqa: E402,F401


class Vec3:
    def __init__(self, x, y, z):
        self.x, self.y, self.z = float(x), float(y), float(z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __getitem__(self, i):
        return (self.x, self.y, self.z)[i]

    def __len__(self):
        return 3

    def __repr__(self):
        return "Vec3(x=%r, y=%r, z=%r)" % (self.x, self.y, self.z)


def _rows(vectors):
    return [[v[0], v[1], v[2]] for v in vectors]


class System(_mm.System):
    def setDefaultPeriodicBoxVectors(self, a, b, c):
        super().setDefaultPeriodicBoxVectors(*_rows((a, b, c)))


class VerletIntegrator:
    def __init__(self, stepSize):
        self.stepSize = stepSize


class NonbondedForce(_mm.SlicedNonbondedForce):
    """OpenMM's NonbondedForce as a parameter container; in a Context it is a SlicedNonbondedForce with one subset."""
    def __init__(self):
        super().__init__(1)


class Platform:
    _names = {"Reference": "Reference", "CUDA": "B200"}

    def __init__(self, name):
        self._name = name

    def getName(self):
        return self._name

    @staticmethod
    def getPlatformByName(name):
        if name not in Platform._names:
            raise Exception('There is no registered Platform called "%s"' % name)
        return Platform(name)


class State:
    def __init__(self, inner, positions, want):
        self._inner, self._positions, self._want = inner, positions, want

    def getPotentialEnergy(self):
        return self._inner.getPotentialEnergy()

    def getForces(self):
        return [Vec3(*f) for f in self._inner.getForces()]

    def getPositions(self):
        return [Vec3(*p) for p in self._positions]

    def getVelocities(self):
        return [Vec3(0, 0, 0) for _ in self._positions]

    def getEnergyParameterDerivatives(self):
        return self._inner.getEnergyParameterDerivatives()


class Context(_mm.Context):
    def __init__(self, system, integrator, platform, properties=None):
        props = {}
        for key, value in (properties or {}).items():
            if key == "Precision":
                props["Precision"] = "double" if value == "double" else "mixed"      # (there is no single-precision mode: mixed)
            else:
                props[key] = value
        super().__init__(system, Platform._names[platform.getName()], props)

    def setPositions(self, positions):
        super().setPositions(np.array(_rows(positions), dtype=np.float64))

    def getState(self, getPositions=False, getVelocities=False, getForces=False, getEnergy=False, getParameters=False,
                 getParameterDerivatives=False, enforcePeriodicBox=False, groups=-1):
        if isinstance(groups, (set, frozenset, list, tuple)):
            mask = 0
            for g in groups:
                mask |= 1 << g
        else:
            mask = groups & 0xFFFFFFFF
        types = (_mm.State.Forces if getForces else 0) | (_mm.State.Energy if getEnergy else 0)
        inner = super().getState(types, enforcePeriodicBox, mask)
        return State(inner, self._positions.copy(), None)
