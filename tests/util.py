"""Shared helpers for the test-suite: the reference's own comparison metric and the
platform matrix (CPU oracle platforms unmarked, CUDA platform marked gpu)."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle_kernel  # noqa: E402

ONE_4PI_EPS0 = 138.93545764438198   # include/snb_constants.h

CPU_PLATFORMS = (["ReferenceRef"] if oracle_kernel.have_ref() else [])+["ReferencePort"]
PLATFORMS = CPU_PLATFORMS+[pytest.param("B200", marks=pytest.mark.gpu)]


def assert_equal_to(expected, found, tol):
    # assertEqualTo, openmmapi/include/internal/AssertionUtilities.h:7-14
    scale = max(abs(expected), 1.0)
    assert abs(expected-found)/scale <= tol, "Expected %r, found %r (rel %g > %g)" % (expected, found, abs(expected-found)/scale, tol)


def assert_equal_vec(expected, found, tol):
    # assertEqualVec, AssertionUtilities.h:16-27
    expected = np.asarray(expected, dtype=float)
    found = np.asarray(found, dtype=float)
    scale = max(math.sqrt(expected.dot(expected)), 1.0)
    err = np.abs(expected-found).max()/scale
    assert err <= tol, "Expected %r, found %r (rel %g > %g)" % (expected, found, err, tol)


def max_force_error(expected, found):
    """max over atoms/components of |dF| / max(|F_ref|, 1)  (assertForces metric)."""
    expected = np.asarray(expected, dtype=float)
    found = np.asarray(found, dtype=float)
    scale = np.maximum(np.sqrt((expected**2).sum(axis=1)), 1.0)
    return (np.abs(expected-found).max(axis=1)/scale).max()


def rel_err(expected, found):
    return abs(expected-found)/max(abs(expected), 1.0)


def evaluate(platform, w, direct=True, reciprocal=True, forces=True, energy=True, parameters=None, pairs=False):
    """One kernel.execute on `platform` for workload `w`; returns energies, forces, derivatives (and pairs)."""
    import openmmlab_b200 as mm
    ctx = mm.Context(w.system, platform)
    ctx.setPositions(w.positions)
    if w.box is not None:
        ctx.setPeriodicBoxVectors(*w.box)
    for k, v in dict(w.parameters, **(parameters or {})).items():
        ctx.setParameter(k, v)
    kernel = ctx._kernels[0][1]
    ctx._forces = np.zeros((w.num_particles, 3))
    ctx._energyParameterDerivatives = {}
    e = kernel.execute(ctx, forces, energy, direct, reciprocal)
    out = {"energy": e, "forces": ctx._forces.copy(), "slices": kernel.sliceEnergies.copy(),
           "derivs": dict(ctx._energyParameterDerivatives), "kernel": kernel, "context": ctx}
    if pairs:
        out["pairs"] = kernel.dumpPairs()
    return out


def assert_parity(ref, got, tol=1e-5, what=""):
    """The reference's own metric (AssertionUtilities.h:7-37) on total energy, every slice energy,
    every derivative and every force component."""
    assert rel_err(ref["energy"], got["energy"]) <= tol, "%s energy %r vs %r" % (what, ref["energy"], got["energy"])
    for s in range(ref["slices"].shape[0]):
        for t in range(2):
            assert rel_err(ref["slices"][s, t], got["slices"][s, t]) <= tol, \
                "%s slice %d term %d: %r vs %r" % (what, s, t, ref["slices"][s, t], got["slices"][s, t])
    assert set(ref["derivs"]) == set(got["derivs"])
    for k in ref["derivs"]:
        assert rel_err(ref["derivs"][k], got["derivs"][k]) <= tol, "%s derivative %s: %r vs %r" % (what, k, ref["derivs"][k], got["derivs"][k])
    err = max_force_error(ref["forces"], got["forces"])
    assert err <= tol, "%s max force error %g > %g" % (what, err, tol)
    return err
