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


# ---- testOpenMMLab (tests/TestSlicedNonbondedForce.h:987-1163) -------------------------------------------------
# The reference compares a SlicedNonbondedForce whose scaling parameters switch one subset off against OpenMM's
# NonbondedForce with the particle / exception parameters scaled by hand.  NonbondedForce is not available offline;
# a ONE-subset SlicedNonbondedForce without scaling parameters is the same function of (q, sigma, eps) and stands in
# for it.  Charges of subset 1 scaled by x scale slice (0,1) by x and slice (1,1) by x^2;
# epsilons scaled by x scale them by sqrt(x) and x (Lorentz-Berthelot) -- hence the parameter pairs below.
def _molecules(method, exceptions, num_subsets):
    num_molecules, cutoff = 100, 3.5
    L = 7.0 if exceptions else 10.0
    system = System()
    for _ in range(2*num_molecules):
        system.addParticle(1.0)
    system.setDefaultPeriodicBoxVectors([L, 0, 0], [0, L, 0], [0, 0, L])
    force = F(num_subsets)
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
        positions[2*k], positions[2*k+1] = center+delta, center-delta
        force.addParticle(1, 1, 1)
        force.addParticle(-1, 1, 1)
        if exceptions:
            force.addException(2*k, 2*k+1, -0.5, 1, 2)
    system.addForce(force)
    return system, force, positions


@pytest.mark.parametrize("platform", PLATFORMS)
@pytest.mark.parametrize("lj", [False, True])
@pytest.mark.parametrize("exceptions", [False, True])
@pytest.mark.parametrize("method", METHODS)
def test_scaling_matches_hand_scaled_parameters(platform, method, exceptions, lj):
    tol = 1e-3 if platform == "B200" else 1e-4
    include_lj, include_coulomb = lj, not lj
    q = lambda k: 1-2*(k % 2)
    qiqj, eps, epsij = -0.5, 1.0, 2.0
    subsets = np.random.default_rng(11).random(200) < 0.5
    system1, plain, positions = _molecules(method, exceptions, 1)
    system2, sliced, _ = _molecules(method, exceptions, 2)
    for k in range(200):
        if subsets[k]:
            sliced.setParticleSubset(k, 1)
    param1 = "lambda" if include_coulomb else "sqrtLambda"
    param2 = "lambdaSq" if include_coulomb else "lambda"
    sliced.addGlobalParameter(param1, 1)
    sliced.addScalingParameter(param1, 0, 1, include_coulomb, include_lj)
    sliced.addGlobalParameter(param2, 1)
    sliced.addScalingParameter(param2, 1, 1, include_coulomb, include_lj)
    particle_scale = [("lambda" if include_coulomb else "one", "lambda" if include_lj else "one") if subsets[k] else ("one", "one")
                      for k in range(200)]
    exception_scale = []
    for k in range(plain.getNumExceptions()):
        si, sj = int(subsets[2*k]), int(subsets[2*k+1])
        if si != sj or si == 1:
            p = param1 if si != sj else param2
            exception_scale.append((p if include_coulomb else "one", p if include_lj else "one"))
        else:
            exception_scale.append(("one", "one"))
    context1, context2 = Context(system1, platform), Context(system2, platform)
    context1.setPositions(positions)
    context2.setPositions(positions)

    def compare():
        for groups in (1 << 0, 1 << 1, 0xFFFFFFFF):
            s1 = context1.getState(State.Energy | State.Forces, False, groups)
            s2 = context2.getState(State.Energy | State.Forces, False, groups)
            assert_equal_to(s1.getPotentialEnergy(), s2.getPotentialEnergy(), tol)
            for f1, f2 in zip(s1.getForces(), s2.getForces()):
                assert_equal_vec(f1, f2, tol)
        return context1.getState(State.Energy).getPotentialEnergy()

    def scale_by(value):
        for k in range(200):
            plain.setParticleParameters(k, q(k)*value[particle_scale[k][0]], 1, eps*value[particle_scale[k][1]])
        for k in range(plain.getNumExceptions()):
            plain.setExceptionParameters(k, 2*k, 2*k+1, qiqj*value[exception_scale[k][0]], 1, epsij*value[exception_scale[k][1]])
        plain.updateParametersInContext(context1)
        context2.setParameter(param1, value[param1])
        context2.setParameter(param2, value[param2])

    energy_one = compare()
    scale_by({"one": 1.0, "lambda": 0.0, "sqrtLambda": 0.0, "lambdaSq": 0.0})
    energy_zero = compare()
    scale_by({"one": 1.0, "lambda": 0.5, "sqrtLambda": 0.5**0.5, "lambdaSq": 0.25})
    compare()
    # derivatives (:1141-1146): dE/dp1 + dE/dp2 = E(1) - E(0) because the scaled energy is linear in both
    sliced.addScalingParameterDerivative(param1)
    sliced.addScalingParameterDerivative(param2)
    context2.reinitialize(True)
    d = context2.getState(State.ParameterDerivatives).getEnergyParameterDerivatives()
    assert_equal_to(energy_one-energy_zero, d[param1]+d[param2], tol)
    # sum of derivatives (:1148-1165): with only the scaled interaction type present, E = sum over slices of dE/dlambda
    for k in range(200):
        plain.setParticleParameters(k, q(k) if include_coulomb else 0, 1, eps if include_lj else 0)
    for k in range(plain.getNumExceptions()):
        plain.setExceptionParameters(k, 2*k, 2*k+1, qiqj if include_coulomb else 0, 1, epsij if include_lj else 0)
    plain.updateParametersInContext(context1)
    energy = context1.getState(State.Energy | State.Forces).getPotentialEnergy()
    sliced.addGlobalParameter("remainder", 1.0)
    sliced.addScalingParameter("remainder", 0, 0, include_coulomb, include_lj)
    sliced.addScalingParameterDerivative("remainder")
    context2.reinitialize(True)
    d = context2.getState(State.Energy | State.ParameterDerivatives).getEnergyParameterDerivatives()
    assert_equal_to(energy, d[param1]+d[param2]+d["remainder"], tol)
