"""The reference's slicing self-consistency test, restated: testScalingParameterSeparation
(tests/TestSlicedNonbondedForce.h:1177-1293), run as the reference runs it -- for every nonbonded method
(NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME, LJPME; :1299-1336) with and without exceptions.

Two SlicedNonbondedForces over the same 100 two-site molecules, two subsets drawn at random:
  force 1: "lambda" scales slice (0,1) Coulomb+LJ, "alpha" slice (0,0), "beta" slice (1,1);
  force 2: "lambdaCoulomb" / "lambdaLJ" scale slice (0,1) separately, "gamma" scales slices (0,0) and (1,1).
With equal values the two must agree in energy and forces, and
  dE/dlambda = dE/dlambdaCoulomb + dE/dlambdaLJ,   E = lambda dE/dlambda + value (dE/dalpha + dE/dbeta),
  dE/dalpha + dE/dbeta = dE/dgamma
-- overall, in direct space alone and in reciprocal space alone.  It needs no NonbondedForce, so it pins the
oracle (CPU platforms) and the CUDA path (platform "B200", marked gpu) on the reference's own terms.
"""
import numpy as np
import pytest

from openmmlab_b200 import Context, SlicedNonbondedForce, State, System
from oracle import oracle_kernel  # noqa: F401  (registers the CPU platforms)

from util import PLATFORMS, assert_equal_to, assert_equal_vec

F = SlicedNonbondedForce
METHODS = [F.NoCutoff, F.CutoffNonPeriodic, F.CutoffPeriodic, F.Ewald, F.PME, F.LJPME]


def _build(method, exceptions, second, rng_subsets):
    num_molecules, cutoff = 100, 3.5
    L = 7.0 if exceptions else 10.0
    system = System()
    for _ in range(2*num_molecules):
        system.addParticle(1.0)
    system.setDefaultPeriodicBoxVectors([L, 0, 0], [0, L, 0], [0, 0, L])
    force = F(2)
    force.setNonbondedMethod(method)
    force.setCutoffDistance(cutoff)
    force.setUseDispersionCorrection(True)
    force.setReciprocalSpaceForceGroup(1)
    force.setEwaldErrorTolerance(1e-4)
    positions = np.zeros((2*num_molecules, 3))
    M = int(round(num_molecules**(1.0/3.0)))
    if M*M*M < num_molecules:
        M += 1
    for k in range(num_molecules):
        iz = k//(M*M)
        iy = (k-iz*M*M)//M
        ix = k-M*(iy+iz*M)
        center = np.array([ix+0.5, iy+0.5, iz+0.5])*L/M
        delta = np.array([0.5-ix % 2, 0.5-iy % 2, 0.5-iz % 2])/2
        i, j = 2*k, 2*k+1
        positions[i], positions[j] = center+delta, center-delta
        force.addParticle(1-2*(i % 2), 1, 1)
        force.addParticle(1-2*(j % 2), 1, 1)
        if exceptions:
            force.addException(i, j, -0.5, 1, 2)
    for k in range(2*num_molecules):
        if rng_subsets[k]:
            force.setParticleSubset(k, 1)
    lam, value = 0.5, 0.3
    if not second:
        force.addGlobalParameter("lambda", lam)
        force.addScalingParameter("lambda", 0, 1, True, True)
        force.addScalingParameterDerivative("lambda")
        force.addGlobalParameter("alpha", value)
        force.addScalingParameter("alpha", 0, 0, True, True)
        force.addScalingParameterDerivative("alpha")
        force.addGlobalParameter("beta", value)
        force.addScalingParameter("beta", 1, 1, True, True)
        force.addScalingParameterDerivative("beta")
    else:
        force.addGlobalParameter("lambdaCoulomb", lam)
        force.addGlobalParameter("lambdaLJ", lam)
        force.addScalingParameter("lambdaCoulomb", 0, 1, True, False)
        force.addScalingParameter("lambdaLJ", 0, 1, False, True)
        force.addScalingParameterDerivative("lambdaCoulomb")
        force.addScalingParameterDerivative("lambdaLJ")
        force.addGlobalParameter("gamma", value)
        force.addScalingParameter("gamma", 0, 0, True, True)
        force.addScalingParameter("gamma", 1, 1, True, True)
        force.addScalingParameterDerivative("gamma")
    system.addForce(force)
    return system, positions, lam, value


@pytest.mark.parametrize("platform", PLATFORMS)
@pytest.mark.parametrize("exceptions", [False, True])
@pytest.mark.parametrize("method", METHODS)
def test_scaling_parameter_separation(platform, method, exceptions):
    # the reference's tolerance: 1e-4 on Reference / double precision, 1e-3 otherwise (:1182)
    tol = 1e-3 if platform == "B200" else 1e-4
    subsets = np.random.default_rng(5).random(200) < 0.5
    system1, positions, lam, value = _build(method, exceptions, False, subsets)
    system2, _, _, _ = _build(method, exceptions, True, subsets)
    context1, context2 = Context(system1, platform), Context(system2, platform)
    context1.setPositions(positions)
    context2.setPositions(positions)
    everything = State.Energy | State.Forces | State.ParameterDerivatives
    for groups in (0xFFFFFFFF, 1 << 0, 1 << 1):     # overall, direct space, reciprocal space (:1252-1292)
        state1 = context1.getState(everything, False, groups)
        state2 = context2.getState(everything, False, groups)
        d1, d2 = state1.getEnergyParameterDerivatives(), state2.getEnergyParameterDerivatives()
        assert_equal_to(state1.getPotentialEnergy(), state2.getPotentialEnergy(), tol)
        for f1, f2 in zip(state1.getForces(), state2.getForces()):
            assert_equal_vec(f1, f2, tol)
        assert_equal_to(d1["lambda"], d2["lambdaCoulomb"]+d2["lambdaLJ"], tol)
        assert_equal_to(state1.getPotentialEnergy(), lam*d1["lambda"]+value*(d1["alpha"]+d1["beta"]), tol)
        assert_equal_to(d1["alpha"]+d1["beta"], d2["gamma"], tol)
