"""Parity of the CUDA path (platform "B200", through the C-ABI) with the CPU oracle on the same
seeded inputs: bit-exact interacting-pair sets, and total energy / per-slice energies / derivatives /
forces within 1e-5 by the reference's own metric (north-star tolerance, mixed precision)."""
import numpy as np
import pytest

import openmmlab_b200 as mm
from openmmlab_b200 import SlicedNonbondedForce as F
from openmmlab_b200 import systems

from util import assert_parity, assert_parity_mixed, evaluate

pytestmark = pytest.mark.gpu
TOL = 1e-5   # BASELINE.json north_star: "within 1e-5 relative error in mixed precision"


DOUBLE = {"Precision": "double"}


def _check_pairs(ref, got):
    assert ref["pairs"].shape == got["pairs"].shape, "pair count %d vs %d" % (len(ref["pairs"]), len(got["pairs"]))
    assert np.array_equal(ref["pairs"], got["pairs"]), "interacting pair sets differ"


def _check(w, pairs=True, mixed=True, energy_tol=1e-4, **kw):
    """Parity of one workload with the CPU oracle in both precision modes of the platform.

    Precision=double: bit-exact pair set; total energy, every slice energy, every derivative and every
    force component within 1e-5 by the reference's own metric -- direct-only, reciprocal-only and both.
    Precision=mixed (default, the fast path): bit-exact pair set; util.assert_parity_mixed.
    """
    method = w.force.getNonbondedMethod()
    direct, reciprocal = kw.pop("direct", True), kw.pop("reciprocal", True)
    ewald_family = method in (F.Ewald, F.PME, F.LJPME)
    want_pairs = pairs and method != F.NoCutoff
    combos = []
    if direct:
        combos.append(("direct", dict(direct=True, reciprocal=False)))
    if reciprocal and ewald_family:
        combos.append(("recip", dict(direct=False, reciprocal=True)))
    if len(combos) == 2:
        combos.append(("total", dict(direct=True, reciprocal=True)))
    refs = {}
    for name, flags in combos:
        ref = refs[name] = evaluate("Reference", w, pairs=want_pairs and name == "direct", **flags, **kw)
        got = evaluate("B200", w, pairs=pairs and name == "direct", properties=DOUBLE, **flags, **kw)
        if "pairs" in ref:
            _check_pairs(ref, got)
        assert_parity(ref, got, TOL, "%s %s [double]" % (w.name, name))
        if mixed:
            # both direct-space kernels of the PME path (DirectKernel; elsewhere the property changes nothing, so once is enough)
            two = direct and method == F.PME and name != "recip"
            for kernel in (("general", "packed") if two else ("auto",)):
                got = evaluate("B200", w, pairs=pairs and name == "direct", properties={"DirectKernel": kernel}, **flags, **kw)
                if "pairs" in ref:
                    _check_pairs(ref, got)
                what = "%s %s [mixed, DirectKernel=%s]" % (w.name, name, kernel)
                if name == "total":
                    assert_parity_mixed(ref, got, refs["direct"], refs["recip"], what=what, energy_tol=energy_tol)
                else:
                    assert_parity_mixed(ref, got, what=what, energy_tol=energy_tol)


@pytest.mark.parametrize("method", [F.CutoffPeriodic, F.PME])
def test_water_box(method):
    # BASELINE.json configs[0]: 648-atom water box, 2 subsets, CutoffPeriodic and PME
    _check(systems.water_box(method=method), energy_tol=1e-5)


@pytest.mark.parametrize("direct,reciprocal", [(True, False), (False, True)])
def test_water_box_direct_reciprocal_split(direct, reciprocal):
    _check(systems.water_box(), direct=direct, reciprocal=reciprocal, energy_tol=1e-5)


@pytest.mark.parametrize("method", [F.CutoffPeriodic, F.Ewald, F.PME, F.LJPME])
@pytest.mark.parametrize("groups", [1, 2, 3])
def test_small_mixed(method, groups):
    kw = {}
    if method == F.LJPME:
        kw["ljpme"] = (systems.ALPHA, 12, 12, 12)
    scaling = [("lam_a", 0.3, 0, 1, True, True), ("lam_b", 0.65, 1, 1, True, False)]
    if groups >= 2:
        scaling.append(("lam_c", 0.9, 0, 2, False, True))
    w = systems.small_mixed(method=method, num_groups=groups, scaling=scaling, derivatives=["lam_a", "lam_b"], **kw)
    _check(w)


def test_switching_function():
    w = systems.small_mixed(method=F.PME, switch=0.75, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"])
    _check(w)
    w = systems.small_mixed(method=F.CutoffPeriodic, switch=0.8, dispersion_correction=True)
    _check(w)


def test_nonperiodic_methods():
    for method in (F.NoCutoff, F.CutoffNonPeriodic):
        w = systems.small_mixed(method=method, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"])
        # A droplet in vacuum: NoCutoff sums 217k unscreened 1/r terms, CutoffNonPeriodic 1e5 reaction-field
        # terms, that cancel to ~1e-5 of their gross size.  In mixed precision these two configurations are
        # held to the reference's own mixed tolerance (1e-3, TestSlicedNonbondedForce.h:995);
        # Precision=double is held to 1e-5 as everywhere else.
        _check(w, energy_tol=1e-3)


def test_triclinic():
    w = systems.small_mixed(method=F.PME, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"])
    L = w.box[0][0]
    box = np.array([[L, 0, 0], [0.25*L, L, 0], [-0.15*L, 0.3*L, L]]).astype(np.float32).astype(np.float64)
    w.box = box
    w.system.setDefaultPeriodicBoxVectors(*box)
    _check(w)


def test_periodic_exceptions_and_many_subsets():
    w = systems.small_mixed(method=F.PME, num_groups=5, scaling=[("lam", 0.5, 2, 4, True, True)], derivatives=["lam"])
    w.force.setExceptionsUsePeriodicBoundaryConditions(True)
    _check(w)


def test_fep_ljpme():
    # BASELINE.json configs[3]
    _check(systems.fep_ligand(), energy_tol=1e-5)


def test_dhfr_size():
    # BASELINE.json configs[1]
    _check(systems.dhfr_like(), energy_tol=1e-5)


def test_lambda_changes_and_offsets():
    w = systems.small_mixed(method=F.PME, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"])
    w.force.addGlobalParameter("off", 0.0)
    w.force.addParticleParameterOffset("off", 3, 0.2, 0.01, 0.1)
    w.force.addParticleParameterOffset("off", 40, -0.2, 0.0, 0.05)
    for lam, off in ((0.0, 0.0), (1.0, 0.5), (0.37, 1.0)):
        _check(w, parameters={"lam": lam, "off": off})


def test_update_parameters():
    # copyParametersToContext (reference test idea: testChangingParameters, TestSlicedNonbondedForce.h:667-732)
    w = systems.water_box()
    results = {}
    for platform, props in (("Reference", None), ("B200", DOUBLE)):
        ctx = mm.Context(w.system, platform, props)
        ctx.setPositions(w.positions)
        ctx.getState()
        for i in range(0, w.num_particles, 5):
            q, s, e = w.force.getParticleParameters(i)
            w.force.setParticleParameters(i, 1.5*q, 1.1*s, 1.7*e)
        w.force.updateParametersInContext(ctx)
        st = ctx.getState(mm.State.Energy | mm.State.Forces | mm.State.ParameterDerivatives)
        results[platform] = {"energy": st.getPotentialEnergy(), "forces": st.getForces(), "derivs": st.getEnergyParameterDerivatives(),
                             "slices": ctx._kernels[0][1].sliceEnergies.copy()}
        for i in range(0, w.num_particles, 5):
            q, s, e = w.force.getParticleParameters(i)
            w.force.setParticleParameters(i, q/1.5, s/1.1, e/1.7)
    assert_parity(results["Reference"], results["B200"], TOL, "updated parameters")


def test_deterministic_forces():
    # fixed-point accumulation: two evaluations give bit-identical forces (reference test idea:
    # platforms/cuda/tests/TestCudaSlicedNonbondedForce.cpp:109-141)
    w = systems.water_box()
    for props in ({"DeterministicForces": "true"}, DOUBLE):
        a = evaluate("B200", w, properties=props)
        b = evaluate("B200", w, properties=props)
        assert np.array_equal(a["forces"], b["forces"])
        assert a["energy"] == b["energy"] or abs(a["energy"]-b["energy"]) < 1e-9*abs(a["energy"])


def test_graph_replay_follows_positions_box_and_lambdas():
    # From the second evaluation of a given shape on, the platform replays a captured CUDA graph; the box,
    # the lambdas and the positions reach it through device memory.  Every replay must still match the oracle.
    w = systems.water_box()
    rng = np.random.default_rng(11)
    contexts = {}
    for platform, props in (("Reference", None), ("B200", DOUBLE), ("B200", None)):
        ctx = mm.Context(w.system, platform, props)
        contexts[(platform, str(props))] = ctx
    base = w.positions.copy()
    L = w.box[0][0]
    replays = 0
    for step in range(5):
        pos = (base+rng.normal(scale=0.004*step, size=base.shape)).astype(np.float32).astype(np.float64)
        scale = np.float64(np.float32(1.0+0.002*step))
        box = np.diag([L*scale]*3)
        lam = [0.8, 0.1, 1.0, 0.55, 0.3][step]
        results = {}
        for key, ctx in contexts.items():
            ctx.setPositions(pos*scale)
            ctx.setPeriodicBoxVectors(*box)
            ctx.setParameter("lambda", lam)
            kernel = ctx._kernels[0][1]
            ctx._forces = np.zeros((w.num_particles, 3))
            ctx._energyParameterDerivatives = {}
            e = kernel.execute(ctx, True, True, True, True)
            results[key] = {"energy": e, "forces": ctx._forces.copy(), "slices": kernel.sliceEnergies.copy(),
                            "derivs": dict(ctx._energyParameterDerivatives)}
            if key[0] == "B200":
                replays += kernel.getCounters()[4]
        ref = results[("Reference", "None")]
        assert_parity(ref, results[("B200", str(DOUBLE))], TOL, "graph replay step %d [double]" % step)
        got = results[("B200", "None")]
        assert abs(ref["energy"]-got["energy"]) <= 1e-5*max(abs(ref["energy"]), np.abs(ref["slices"]).sum())
        assert np.abs(ref["forces"]-got["forces"]).max() <= 1e-4*np.sqrt((ref["forces"]**2).sum(axis=1).mean())
    assert replays >= 6, "the CUDA-graph path was not exercised (%d replays)" % replays


def test_apoa1_size():
    # BASELINE.json configs[2]: 92,224 atoms, 4 subsets / 10 slices, three scaling parameters with derivatives.
    # Oracle: direct and reciprocal space separately and together; Precision=double at 1e-5, mixed by its criteria.
    _check(systems.apoa1_like(), energy_tol=1e-5)
