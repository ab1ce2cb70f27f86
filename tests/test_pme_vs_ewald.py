"""PME pinned against an independent method: classic Ewald summation on the 648-atom water box.

The reference checks its PME only against OpenMM's NonbondedForce, which is not available offline.  Its Ewald branch
(ReferenceSlicedLJCoulombIxn.cpp:240-342) shares nothing with the PME branch beyond the direct-space loop -- no
B-splines, no grid, no FFT -- so with both driven to a tight tolerance the two must give the same total energy, the
same per-slice energies and the same forces.  Cutoff and alpha are the same in both runs (alpha follows from cutoff and
tolerance), so direct space cancels exactly and the comparison is reciprocal space against reciprocal space.
"""
import numpy as np
import pytest

from openmmlab_b200 import SlicedNonbondedForce as F, systems
from oracle import oracle_kernel  # noqa: F401  (registers the CPU platforms)

from util import PLATFORMS, evaluate


def _water(method, tolerance):
    w = systems.water_box(method=method)
    w.force.setEwaldErrorTolerance(tolerance)
    w.force.setPMEParameters(0.0, 0, 0, 0)          # alpha and grid from the tolerance, as for Ewald's alpha and kmax
    return w


@pytest.mark.parametrize("platform", PLATFORMS)
def test_pme_agrees_with_ewald_summation(platform):
    tolerance = 1e-6
    pme = evaluate(platform, _water(F.PME, tolerance))
    ewald = evaluate(platform, _water(F.Ewald, tolerance))
    # both methods carry their own truncation error of order `tolerance` relative to the force / energy scale
    tol = 2e-5 if platform != "B200" else 1e-4
    scale = np.abs(ewald["slices"]).max()
    assert abs(pme["energy"]-ewald["energy"]) <= tol*scale
    assert np.abs(pme["slices"]-ewald["slices"]).max() <= tol*scale
    for name in ewald["derivs"]:
        assert abs(pme["derivs"][name]-ewald["derivs"][name]) <= tol*scale
    rms = np.sqrt((ewald["forces"]**2).sum(axis=1).mean())
    assert np.abs(pme["forces"]-ewald["forces"]).max() <= 10*tol*rms
