"""Tile-list reuse (platform property ListSkin, include/slicednb.h snb_set_list_skin).

With a skin the sorted order and the tile list are built for cutoff + skin and kept until an atom has moved more than
skin/2; the exact cutoff predicate still decides the pair set every evaluation.  So along any trajectory a context with a
skin must give the SAME set of interacting pairs as one that rebuilds every evaluation (the Reference platform's behaviour,
platforms/reference/src/ReferenceOpenMMLabKernels.cpp:203) and energies / forces / derivatives that agree to the
precision of the mode -- while rebuilding only now and then.  (OpenMM's CUDA platform, whose neighbour list the reference's
CUDA kernels inherit, pads and reuses its list the same way.)"""
import numpy as np
import pytest

import openmmlab_b200 as mm
from openmmlab_b200 import systems

pytestmark = pytest.mark.gpu


def _context(w, props):
    ctx = mm.Context(w.system, "B200", props)
    if w.box is not None:
        ctx.setPeriodicBoxVectors(*w.box)
    for k, v in w.parameters.items():
        ctx.setParameter(k, v)
    return ctx


def _step(ctx, positions, pairs=True):
    kernel = ctx._kernels[0][1]
    ctx.setPositions(positions)
    ctx._forces = np.zeros((len(positions), 3))
    ctx._energyParameterDerivatives = {}
    e = kernel.execute(ctx, True, True, True, True)
    out = {"energy": e, "forces": ctx._forces.copy(), "slices": kernel.sliceEnergies.copy(), "derivs": dict(ctx._energyParameterDerivatives)}
    if pairs:
        out["pairs"] = kernel.dumpPairs()
    return out


def _compare(ref, got, tol, what, truth=None):
    """`got` (kept list) against `ref` (list rebuilt every evaluation).  The pair sets must be EQUAL.  Energies to `tol` of
    the gross energy.  Forces: in mixed precision both carry the FP32 rounding of their own block-local coordinates, so
    they are held against `truth` (Precision=double, list rebuilt every evaluation): the kept list may not be further from
    it than three times what the rebuilt one is (+ 1e-6 of the rms force); without `truth`, directly to 10 tol."""
    assert np.array_equal(ref["pairs"], got["pairs"]), "%s: pair sets differ (%d vs %d)" % (what, len(ref["pairs"]), len(got["pairs"]))
    scale = max(1.0, np.abs(ref["slices"]).sum())
    assert abs(ref["energy"]-got["energy"]) <= tol*scale, "%s: energy %r vs %r" % (what, ref["energy"], got["energy"])
    assert np.abs(ref["slices"]-got["slices"]).max() <= tol*scale, what
    for k in ref["derivs"]:
        assert abs(ref["derivs"][k]-got["derivs"][k]) <= tol*scale, "%s: derivative %s" % (what, k)
    rms = np.sqrt((ref["forces"]**2).sum(axis=1).mean())
    if truth is None:
        err = np.abs(ref["forces"]-got["forces"]).max()
        assert err <= 10*tol*rms, "%s: forces differ by %g of the rms force" % (what, err/rms)
    else:
        assert np.array_equal(truth["pairs"], got["pairs"]), "%s: pair set differs from the double-precision one" % what
        e_ref, e_got = np.abs(ref["forces"]-truth["forces"]).max(), np.abs(got["forces"]-truth["forces"]).max()
        assert e_got <= 3*e_ref+1e-6*rms, "%s: force error %g of the rms force with the kept list, %g with a fresh one" % (what, e_got/rms, e_ref/rms)


CASES = [("small PME general", lambda: systems.small_mixed(n_side=10, solute_sites=20, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"]), {"DirectKernel": "general"}, 2e-6),
         ("small PME packed", lambda: systems.small_mixed(n_side=10, solute_sites=20, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"]), {"DirectKernel": "packed"}, 2e-6),
         ("small PME double", lambda: systems.small_mixed(n_side=10, solute_sites=20, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"]), {"Precision": "double"}, 1e-10),
         ("fep LJPME", systems.fep_ligand, {}, 2e-6),
         ("dhfr PME", systems.dhfr_like, {}, 2e-6)]


@pytest.mark.parametrize("name,make,props,tol", CASES, ids=[c[0] for c in CASES])
def test_trajectory_matches_rebuild_every_step(name, make, props, tol):
    w = make()
    rng = np.random.default_rng(11)
    # nm per step: the fastest atoms pass skin/2 after 5-6 steps (much faster, and atoms end up on top of each other within
    # the test: forces beyond the 2^31 kJ/mol/nm range of the fixed-point accumulators, where every path saturates differently)
    velocity = rng.normal(0.0, 0.002, size=w.positions.shape)
    steps = 14 if w.num_particles < 10000 else 8
    fresh = _context(w, dict(props))
    kept = _context(w, dict(props, ListSkin="0.1"))
    truth = _context(w, {"Precision": "double"}) if props.get("Precision") != "double" else None
    for k in range(steps):
        pos = w.positions+k*velocity
        _compare(_step(fresh, pos), _step(kept, pos), tol, "%s step %d" % (name, k), _step(truth, pos) if truth else None)
    counters = kept._kernels[0][1].getCounters()
    rebuilds, evals = counters[6], counters[7]
    assert evals == steps
    assert 2 <= rebuilds < steps, "%d rebuilds in %d evaluations" % (rebuilds, steps)
    assert fresh._kernels[0][1].getCounters()[6] == 0


def test_atoms_crossing_the_box_faces_keep_their_blocks():
    """A rigid drift takes atoms across the faces of the unit cell while nothing moves far enough for a rebuild."""
    w = systems.small_mixed(n_side=10, solute_sites=20)
    fresh = _context(w, {})
    kept = _context(w, {"ListSkin": "0.3"})
    truth = _context(w, {"Precision": "double"})
    drift = np.array([0.011, -0.007, 0.005])
    ref0 = None
    for k in range(10):
        pos = w.positions+k*drift
        a, b = _step(fresh, pos), _step(kept, pos)
        _compare(a, b, 2e-6, "drift step %d" % k, _step(truth, pos))
        if ref0 is None:
            ref0 = a
        # (a rigid translation leaves everything but the last bits alone: the pair set is that of step 0)
        assert np.array_equal(ref0["pairs"], b["pairs"])
    assert kept._kernels[0][1].getCounters()[6] == 1, "a drift of 0.014 nm per step must not rebuild a list with a 0.3 nm skin within 10 steps"


def test_box_change_and_new_skin_rebuild():
    w = systems.small_mixed(n_side=10, solute_sites=20)
    kept = _context(w, {"ListSkin": "0.1"})
    fresh = _context(w, {})
    kernel = kept._kernels[0][1]
    _step(kept, w.positions)
    _step(kept, w.positions)
    assert kernel.getCounters()[6] == 1
    # isotropic scaling of box and positions (a barostat move)
    s = 1.01
    box = [tuple(s*np.asarray(v)) for v in w.box]
    for ctx in (kept, fresh):
        ctx.setPeriodicBoxVectors(*box)
    _compare(_step(fresh, s*w.positions), _step(kept, s*w.positions), 5e-6, "scaled box")
    assert kernel.getCounters()[6] == 2
    kernel.setListSkin(0.05)
    _compare(_step(fresh, s*w.positions), _step(kept, s*w.positions), 5e-6, "new skin")
    assert kernel.getCounters()[6] == 3
    kernel.setListSkin(0.0)
    _compare(_step(fresh, s*w.positions), _step(kept, s*w.positions), 5e-6, "skin off")
    assert kernel.getCounters()[6] == 3


def test_skin_is_clipped_in_a_small_box():
    """648 waters in a 1.86 nm box with a 0.9 nm cutoff: 0.03 nm of room; a 0.2 nm skin must not break anything."""
    w = systems.water_box()
    fresh, kept = _context(w, {}), _context(w, {"ListSkin": "0.2"})
    rng = np.random.default_rng(5)
    velocity = rng.normal(0.0, 0.002, size=w.positions.shape)
    for k in range(6):
        _compare(_step(fresh, w.positions+k*velocity), _step(kept, w.positions+k*velocity), 5e-6, "water step %d" % k)


def test_illegal_skin():
    w = systems.water_box()
    with pytest.raises(mm.OpenMMException):
        mm.Context(w.system, "B200", {"ListSkin": "-1"})
