"""Known-answer tests of the reference (tests/TestSlicedNonbondedForce.h), restated one for one.

Each test names the reference test it follows.  They run
  * on the CPU oracle -- "ReferenceRef" (the reference's own sources compiled unchanged, when
    oracle/_ref is built) and "ReferencePort" (the restatement) -- which PINS the oracle to the
    reference's golden answers, and
  * on platform "B200" (the CUDA path through the C-ABI), marked gpu.
Tolerances are the reference's own (TOL = 1e-4, TestSlicedNonbondedForce.h:28).
"""
import math

import numpy as np
import pytest

import openmmlab_b200 as mm
from openmmlab_b200 import Context, SlicedNonbondedForce, State, System
from oracle import oracle_kernel

from util import ONE_4PI_EPS0, assert_equal_to, assert_equal_vec, PLATFORMS

TOL = 1e-4
SQRT_TWO = math.sqrt(2.0)

pytestmark = pytest.mark.parametrize("platform", PLATFORMS)


def test_coulomb(platform):
    # testCoulomb, TestSlicedNonbondedForce.h:73-95
    system = System()
    system.addParticle(1.0)
    system.addParticle(1.0)
    ff = SlicedNonbondedForce(1)
    ff.addParticle(0.5, 1, 0)
    ff.addParticle(-1.5, 1, 0)
    system.addForce(ff)
    assert not ff.usesPeriodicBoundaryConditions()
    assert not system.usesPeriodicBoundaryConditions()
    context = Context(system, platform)
    context.setPositions([[0, 0, 0], [2, 0, 0]])
    state = context.getState(State.Forces | State.Energy)
    forces = state.getForces()
    force = ONE_4PI_EPS0*(-0.75)/4.0
    assert_equal_vec([-force, 0, 0], forces[0], TOL)
    assert_equal_vec([force, 0, 0], forces[1], TOL)
    assert_equal_to(ONE_4PI_EPS0*(-0.75)/2.0, state.getPotentialEnergy(), TOL)


def test_lj(platform):
    # testLJ, :97-121
    system = System()
    system.addParticle(1.0)
    system.addParticle(1.0)
    ff = SlicedNonbondedForce(1)
    ff.addParticle(0, 1.2, 1)
    ff.addParticle(0, 1.4, 2)
    system.addForce(ff)
    context = Context(system, platform)
    context.setPositions([[0, 0, 0], [2, 0, 0]])
    state = context.getState(State.Forces | State.Energy)
    forces = state.getForces()
    x = 1.3/2.0
    eps = SQRT_TWO
    force = 4.0*eps*(12*x**12-6*x**6)/2.0
    assert_equal_vec([-force, 0, 0], forces[0], TOL)
    assert_equal_vec([force, 0, 0], forces[1], TOL)
    assert_equal_to(4.0*eps*(x**12-x**6), state.getPotentialEnergy(), TOL)


def _find14s(sliced):
    first14 = second14 = None
    for i in range(sliced.getNumExceptions()):
        p1, p2 = sliced.getExceptionParameters(i)[:2]
        if {p1, p2} == {0, 3}:
            first14 = i
        if {p1, p2} == {1, 4}:
            second14 = i
    return first14, second14


def test_exclusions_and_14(platform):
    # testExclusionsAnd14, :123-208
    system = System()
    sliced = SlicedNonbondedForce(1)
    for i in range(5):
        system.addParticle(1.0)
        sliced.addParticle(0, 1.5, 0)
    sliced.createExceptionsFromBonds([(0, 1), (1, 2), (2, 3), (3, 4)], 0.0, 0.0)
    first14, second14 = _find14s(sliced)
    system.addForce(sliced)
    context = Context(system, platform)
    for i in range(1, 5):
        # LJ forces
        positions = np.zeros((5, 3))
        r = 1.0
        for j in range(5):
            sliced.setParticleParameters(j, 0, 1.5, 0)
            positions[j] = [0, j, 0]
        sliced.setParticleParameters(0, 0, 1.5, 1)
        sliced.setParticleParameters(i, 0, 1.5, 1)
        sliced.setExceptionParameters(first14, 0, 3, 0, 1.5, 0.5 if i == 3 else 0.0)
        sliced.setExceptionParameters(second14, 1, 4, 0, 1.5, 0.0)
        positions[i] = [r, 0, 0]
        context.reinitialize()
        context.setPositions(positions)
        state = context.getState(State.Forces | State.Energy)
        forces = state.getForces()
        x = 1.5/r
        eps = 1.0
        force = 4.0*eps*(12*x**12-6*x**6)/r
        energy = 4.0*eps*(x**12-x**6)
        if i == 3:
            force *= 0.5
            energy *= 0.5
        if i < 3:
            force = energy = 0
        assert_equal_vec([-force, 0, 0], forces[0], TOL)
        assert_equal_vec([force, 0, 0], forces[i], TOL)
        assert_equal_to(energy, state.getPotentialEnergy(), TOL)
        # Coulomb forces
        sliced.setParticleParameters(0, 2, 1.5, 0)
        sliced.setParticleParameters(i, 2, 1.5, 0)
        sliced.setExceptionParameters(first14, 0, 3, 4/1.2 if i == 3 else 0, 1.5, 0)
        sliced.setExceptionParameters(second14, 1, 4, 0, 1.5, 0)
        context.reinitialize()
        context.setPositions(positions)
        state = context.getState(State.Forces | State.Energy)
        forces2 = state.getForces()
        force = ONE_4PI_EPS0*4/(r*r)
        energy = ONE_4PI_EPS0*4/r
        if i == 3:
            force /= 1.2
            energy /= 1.2
        if i < 3:
            force = energy = 0
        assert_equal_vec([-force, 0, 0], forces2[0], TOL)
        assert_equal_vec([force, 0, 0], forces2[i], TOL)
        assert_equal_to(energy, state.getPotentialEnergy(), TOL)


def test_cutoff(platform):
    # testCutoff (reaction field), :210-246
    system = System()
    for _ in range(3):
        system.addParticle(1.0)
    ff = SlicedNonbondedForce(1)
    for _ in range(3):
        ff.addParticle(1.0, 1, 0)
    ff.setNonbondedMethod(SlicedNonbondedForce.CutoffNonPeriodic)
    cutoff = 2.9
    ff.setCutoffDistance(cutoff)
    eps = 50.0
    ff.setReactionFieldDielectric(eps)
    system.addForce(ff)
    assert not ff.usesPeriodicBoundaryConditions()
    context = Context(system, platform)
    context.setPositions([[0, 0, 0], [0, 2, 0], [0, 3, 0]])
    state = context.getState(State.Forces | State.Energy)
    forces = state.getForces()
    krf = (1.0/cutoff**3)*(eps-1.0)/(2.0*eps+1.0)
    crf = (1.0/cutoff)*(3.0*eps)/(2.0*eps+1.0)
    force1 = ONE_4PI_EPS0*(0.25-2.0*krf*2.0)
    force2 = ONE_4PI_EPS0*(1.0-2.0*krf*1.0)
    assert_equal_vec([0, -force1, 0], forces[0], TOL)
    assert_equal_vec([0, force1-force2, 0], forces[1], TOL)
    assert_equal_vec([0, force2, 0], forces[2], TOL)
    energy1 = ONE_4PI_EPS0*(0.5+krf*4.0-crf)
    energy2 = ONE_4PI_EPS0*(1.0+krf*1.0-crf)
    assert_equal_to(energy1+energy2, state.getPotentialEnergy(), TOL)


def test_cutoff_14(platform):
    # testCutoff14, :248-342
    system = System()
    sliced = SlicedNonbondedForce(1)
    sliced.setNonbondedMethod(SlicedNonbondedForce.CutoffNonPeriodic)
    for i in range(5):
        system.addParticle(1.0)
        sliced.addParticle(0, 1.5, 0)
    cutoff = 3.5
    sliced.setCutoffDistance(cutoff)
    sliced.setReactionFieldDielectric(30.0)
    sliced.createExceptionsFromBonds([(0, 1), (1, 2), (2, 3), (3, 4)], 0.0, 0.0)
    first14, second14 = _find14s(sliced)
    system.addForce(sliced)
    context = Context(system, platform)
    positions = np.array([[i, 0, 0] for i in range(5)], dtype=float)
    context.setPositions(positions)
    for i in range(1, 5):
        sliced.setParticleParameters(0, 0, 1.5, 1)
        for j in range(1, 5):
            sliced.setParticleParameters(j, 0, 1.5, 0)
        sliced.setParticleParameters(i, 0, 1.5, 1)
        sliced.setExceptionParameters(first14, 0, 3, 0, 1.5, 0.5 if i == 3 else 0.0)
        sliced.setExceptionParameters(second14, 1, 4, 0, 1.5, 0.0)
        context.reinitialize(True)
        state = context.getState(State.Forces | State.Energy)
        forces = state.getForces()
        r = positions[i][0]
        x = 1.5/r
        force = 4.0*(12*x**12-6*x**6)/r
        energy = 4.0*(x**12-x**6)
        if i == 3:
            force *= 0.5
            energy *= 0.5
        if i < 3 or r > cutoff:
            force = energy = 0
        assert_equal_vec([-force, 0, 0], forces[0], TOL)
        assert_equal_vec([force, 0, 0], forces[i], TOL)
        assert_equal_to(energy, state.getPotentialEnergy(), TOL)
        q = 0.7
        sliced.setParticleParameters(0, q, 1.5, 0)
        sliced.setParticleParameters(i, q, 1.5, 0)
        sliced.setExceptionParameters(first14, 0, 3, q*q/1.2 if i == 3 else 0, 1.5, 0)
        sliced.setExceptionParameters(second14, 1, 4, 0, 1.5, 0)
        context.reinitialize(True)
        state = context.getState(State.Forces | State.Energy)
        forces2 = state.getForces()
        force = ONE_4PI_EPS0*q*q/(r*r)
        energy = ONE_4PI_EPS0*q*q/r
        if i == 3:
            force /= 1.2
            energy /= 1.2
        if i < 3 or r > cutoff:
            force = energy = 0
        assert_equal_vec([-force, 0, 0], forces2[0], TOL)
        assert_equal_vec([force, 0, 0], forces2[i], TOL)
        assert_equal_to(energy, state.getPotentialEnergy(), TOL)


def test_periodic(platform):
    # testPeriodic, :344-378
    system = System()
    for _ in range(3):
        system.addParticle(1.0)
    sliced = SlicedNonbondedForce(1)
    for _ in range(3):
        sliced.addParticle(1.0, 1, 0)
    sliced.addException(0, 1, 0.0, 1.0, 0.0)
    sliced.setNonbondedMethod(SlicedNonbondedForce.CutoffPeriodic)
    cutoff = 2.0
    sliced.setCutoffDistance(cutoff)
    system.setDefaultPeriodicBoxVectors([4, 0, 0], [0, 4, 0], [0, 0, 4])
    system.addForce(sliced)
    assert sliced.usesPeriodicBoundaryConditions()
    assert system.usesPeriodicBoundaryConditions()
    context = Context(system, platform)
    context.setPositions([[0, 0, 0], [2, 0, 0], [3, 0, 0]])
    state = context.getState(State.Forces | State.Energy)
    forces = state.getForces()
    eps = 78.3
    krf = (1.0/cutoff**3)*(eps-1.0)/(2.0*eps+1.0)
    crf = (1.0/cutoff)*(3.0*eps)/(2.0*eps+1.0)
    force = ONE_4PI_EPS0*(1.0-2.0*krf*1.0)
    assert_equal_vec([force, 0, 0], forces[0], TOL)
    assert_equal_vec([-force, 0, 0], forces[1], TOL)
    assert_equal_vec([0, 0, 0], forces[2], TOL)
    assert_equal_to(2*ONE_4PI_EPS0*(1.0+krf*1.0-crf), state.getPotentialEnergy(), TOL)


def test_periodic_exceptions(platform):
    # testPeriodicExceptions, :380-416
    system = System()
    system.addParticle(1.0)
    system.addParticle(1.0)
    sliced = SlicedNonbondedForce(1)
    sliced.addParticle(1.0, 1, 0)
    sliced.addParticle(1.0, 1, 0)
    sliced.addException(0, 1, 1.0, 1.0, 0.0)
    sliced.setNonbondedMethod(SlicedNonbondedForce.CutoffPeriodic)
    sliced.setCutoffDistance(2.0)
    system.setDefaultPeriodicBoxVectors([4, 0, 0], [0, 4, 0], [0, 0, 4])
    system.addForce(sliced)
    context = Context(system, platform)
    context.setPositions([[0, 0, 0], [3, 0, 0]])
    state = context.getState(State.Forces | State.Energy)
    forces = state.getForces()
    force = ONE_4PI_EPS0/(3*3)
    assert_equal_vec([-force, 0, 0], forces[0], TOL)
    assert_equal_vec([force, 0, 0], forces[1], TOL)
    assert_equal_to(ONE_4PI_EPS0/3, state.getPotentialEnergy(), TOL)
    sliced.setExceptionsUsePeriodicBoundaryConditions(True)
    context.reinitialize(True)
    state = context.getState(State.Forces | State.Energy)
    forces = state.getForces()
    force = ONE_4PI_EPS0/(1*1)
    assert_equal_vec([force, 0, 0], forces[0], TOL)
    assert_equal_vec([-force, 0, 0], forces[1], TOL)
    assert_equal_to(ONE_4PI_EPS0/1, state.getPotentialEnergy(), TOL)


def test_triclinic(platform):
    # testTriclinic, :418-478 (numpy RNG instead of SFMT; the check is a brute-force image search)
    system = System()
    system.addParticle(1.0)
    system.addParticle(1.0)
    a, b, c = np.array([3.1, 0, 0]), np.array([0.4, 3.5, 0]), np.array([-0.1, -0.5, 4.0])
    system.setDefaultPeriodicBoxVectors(a, b, c)
    sliced = SlicedNonbondedForce(1)
    sliced.addParticle(1.0, 1, 0)
    sliced.addParticle(1.0, 1, 0)
    sliced.setNonbondedMethod(SlicedNonbondedForce.CutoffPeriodic)
    cutoff = 1.5
    sliced.setCutoffDistance(cutoff)
    system.addForce(sliced)
    context = Context(system, platform)
    rng = np.random.default_rng(0)
    eps = 78.3
    krf = (1.0/cutoff**3)*(eps-1.0)/(2.0*eps+1.0)
    crf = (1.0/cutoff)*(3.0*eps)/(2.0*eps+1.0)
    for _ in range(50):
        u = rng.random((2, 3))
        # float-representable inputs so every platform sees identical coordinates
        positions = np.array([a*u[k, 0]+b*u[k, 1]+c*u[k, 2] for k in range(2)]).astype(np.float32).astype(np.float64)
        context.setPositions(positions)
        delta, distance2 = None, 100.0
        for i in (-1, 0, 1):
            for j in (-1, 0, 1):
                for k in (-1, 0, 1):
                    d = positions[1]-positions[0]+a*i+b*j+c*k
                    if d.dot(d) < distance2:
                        delta, distance2 = d, d.dot(d)
        distance = math.sqrt(distance2)
        if abs(distance-cutoff) < 1e-4:
            continue
        state = context.getState(State.Forces | State.Energy)
        if distance >= cutoff:
            assert state.getPotentialEnergy() == 0.0
            assert_equal_vec([0, 0, 0], state.getForces()[0], 0)
            assert_equal_vec([0, 0, 0], state.getForces()[1], 0)
        else:
            force = delta*ONE_4PI_EPS0*(-1.0/(distance**3)+2.0*krf)
            assert_equal_to(ONE_4PI_EPS0*(1.0/distance+krf*distance*distance-crf), state.getPotentialEnergy(), 1e-4)
            assert_equal_vec(force, state.getForces()[0], 1e-4)
            assert_equal_vec(-force, state.getForces()[1], 1e-4)


def test_dispersion_correction(platform):
    # testDispersionCorrection, :600-665
    gridSize = 5
    numParticles = gridSize**3
    boxSize = gridSize*0.7
    cutoff = boxSize/3
    tol = 1e-4
    system = System()
    sliced = SlicedNonbondedForce(1)
    positions = []
    for i in range(gridSize):
        for j in range(gridSize):
            for k in range(gridSize):
                system.addParticle(1.0)
                sliced.addParticle(0, 1.1, 0.5)
                positions.append([i*boxSize/gridSize, j*boxSize/gridSize, k*boxSize/gridSize])
    sliced.setNonbondedMethod(SlicedNonbondedForce.CutoffPeriodic)
    sliced.setCutoffDistance(cutoff)
    system.setDefaultPeriodicBoxVectors([boxSize, 0, 0], [0, boxSize, 0], [0, 0, boxSize])
    system.addForce(sliced)
    context = Context(system, platform)
    context.setPositions(positions)
    energy1 = context.getState(State.Energy).getPotentialEnergy()
    sliced.setUseDispersionCorrection(False)
    context.reinitialize()
    context.setPositions(positions)
    energy2 = context.getState(State.Energy).getPotentialEnergy()
    term1 = (0.5*1.1**12/cutoff**9)/9
    term2 = (0.5*1.1**6/cutoff**3)/3
    expected = 8*math.pi*numParticles*numParticles*(term1-term2)/boxSize**3
    assert_equal_to(expected, energy1-energy2, tol)
    numType2 = 0
    for i in range(0, numParticles, 2):
        sliced.setParticleParameters(i, 0, 1, 1)
        numType2 += 1
    numType1 = numParticles-numType2
    sliced.updateParametersInContext(context)
    energy2 = context.getState(State.Energy).getPotentialEnergy()
    sliced.setUseDispersionCorrection(True)
    context.reinitialize()
    context.setPositions(positions)
    energy1 = context.getState(State.Energy).getPotentialEnergy()
    term1 = ((numType1*(numType1+1))//2)*(0.5*1.1**12/cutoff**9)/9
    term2 = ((numType1*(numType1+1))//2)*(0.5*1.1**6/cutoff**3)/3
    term1 += ((numType2*(numType2+1))//2)*(1.0/cutoff**9)/9
    term2 += ((numType2*(numType2+1))//2)*(1.0/cutoff**3)/3
    combinedSigma = 0.5*(1+1.1)
    combinedEpsilon = math.sqrt(1*0.5)
    term1 += (numType1*numType2)*(combinedEpsilon*combinedSigma**12/cutoff**9)/9
    term2 += (numType1*numType2)*(combinedEpsilon*combinedSigma**6/cutoff**3)/3
    term1 /= (numParticles*(numParticles+1))//2
    term2 /= (numParticles*(numParticles+1))//2
    expected = 8*math.pi*numParticles*numParticles*(term1-term2)/boxSize**3
    assert_equal_to(expected, energy1-energy2, tol)


@pytest.mark.parametrize("method", [SlicedNonbondedForce.CutoffNonPeriodic, SlicedNonbondedForce.PME])
def test_switching_function(platform, method):
    # testSwitchingFunction, :734-787
    system = System()
    system.setDefaultPeriodicBoxVectors([6, 0, 0], [0, 6, 0], [0, 0, 6])
    system.addParticle(1.0)
    system.addParticle(1.0)
    sliced = SlicedNonbondedForce(1)
    sliced.addParticle(0, 1.2, 1)
    sliced.addParticle(0, 1.4, 2)
    sliced.setNonbondedMethod(method)
    sliced.setCutoffDistance(2.0)
    sliced.setUseSwitchingFunction(True)
    sliced.setSwitchingDistance(1.5)
    sliced.setUseDispersionCorrection(False)
    system.addForce(sliced)
    context = Context(system, platform)
    eps = SQRT_TWO
    r = 1.0
    while r < 2.5:
        # the reference steps r by 0.1 in double; r = 2.0 lands a hair below 2.0 there as well
        context.setPositions([[0, 0, 0], [r, 0, 0]])
        state = context.getState(State.Forces | State.Energy)
        x = 1.3/r
        expectedEnergy = 4.0*eps*(x**12-x**6)
        if r <= 1.5:
            switchValue = 1
        elif r >= 2.0:
            switchValue = 0
        else:
            t = (r-1.5)/0.5
            switchValue = 1+t*t*t*(-10+t*(15-t*6))
        assert_equal_to(switchValue*expectedEnergy, state.getPotentialEnergy(), TOL)
        delta = 1e-3
        context.setPositions([[0, 0, 0], [r-delta, 0, 0]])
        e1 = context.getState(State.Energy).getPotentialEnergy()
        context.setPositions([[0, 0, 0], [r+delta, 0, 0]])
        e2 = context.getState(State.Energy).getPotentialEnergy()
        assert_equal_to((e2-e1)/(2*delta), state.getForces()[0][0], 1e-3)
        r += 0.1


def test_two_forces(platform):
    # testTwoForces, :789-837
    system = System()
    system.addParticle(1.0)
    system.addParticle(1.0)
    nb1 = SlicedNonbondedForce(1)
    nb1.addParticle(-1.5, 1, 1.2)
    nb1.addParticle(0.5, 1, 1.0)
    system.addForce(nb1)
    nb2 = SlicedNonbondedForce(1)
    nb2.addParticle(0.4, 1.4, 0.5)
    nb2.addParticle(0.3, 1.8, 1.0)
    nb2.setForceGroup(1)
    system.addForce(nb2)
    context = Context(system, platform)
    context.setPositions([[0, 0, 0], [1.5, 0, 0]])
    state1 = context.getState(State.Energy, False, 1 << 0)
    assert_equal_to(ONE_4PI_EPS0*(-1.5*0.5)/1.5+4.0*math.sqrt(1.2*1.0)*((1.0/1.5)**12-(1.0/1.5)**6), state1.getPotentialEnergy(), TOL)
    state2 = context.getState(State.Energy, False, 1 << 1)
    assert_equal_to(ONE_4PI_EPS0*(0.4*0.3)/1.5+4.0*math.sqrt(0.5*1.0)*((1.6/1.5)**12-(1.6/1.5)**6), state2.getPotentialEnergy(), TOL)
    state = context.getState(State.Energy)
    assert_equal_to(state1.getPotentialEnergy()+state2.getPotentialEnergy(), state.getPotentialEnergy(), TOL)
    nb1.setParticleParameters(0, -1.2, 1.1, 1.4)
    nb1.updateParametersInContext(context)
    nb2.setParticleParameters(0, 0.5, 1.6, 0.6)
    nb2.updateParametersInContext(context)
    state1 = context.getState(State.Energy, False, 1 << 0)
    assert_equal_to(ONE_4PI_EPS0*(-1.2*0.5)/1.5+4.0*math.sqrt(1.4*1.0)*((1.05/1.5)**12-(1.05/1.5)**6), state1.getPotentialEnergy(), TOL)
    state2 = context.getState(State.Energy, False, 1 << 1)
    assert_equal_to(ONE_4PI_EPS0*(0.5*0.3)/1.5+4.0*math.sqrt(0.6*1.0)*((1.7/1.5)**12-(1.7/1.5)**6), state2.getPotentialEnergy(), TOL)
    # PME (default box 2x2x2 -> shrink the cutoff below half the box, as OpenMM's default box requires)
    nb1.setNonbondedMethod(SlicedNonbondedForce.PME)
    nb2.setNonbondedMethod(SlicedNonbondedForce.PME)
    context.reinitialize(True)
    state1 = context.getState(State.Energy, False, 1 << 0)
    state2 = context.getState(State.Energy, False, 1 << 1)
    state = context.getState(State.Energy)
    assert_equal_to(state1.getPotentialEnergy()+state2.getPotentialEnergy(), state.getPotentialEnergy(), TOL)


def test_parameter_offsets(platform):
    # testParameterOffsets, :839-901
    system = System()
    for _ in range(4):
        system.addParticle(1.0)
    force = SlicedNonbondedForce(1)
    force.addParticle(0.0, 1.0, 0.5)
    force.addParticle(1.0, 0.5, 0.6)
    force.addParticle(-1.0, 2.0, 0.7)
    force.addParticle(0.5, 2.0, 0.8)
    force.addException(0, 3, 0.0, 1.0, 0.0)
    force.addException(2, 3, 0.5, 1.0, 1.5)
    force.addException(0, 1, 1.0, 1.5, 1.0)
    force.addGlobalParameter("p1", 0.0)
    force.addGlobalParameter("p2", 1.0)
    force.addParticleParameterOffset("p1", 0, 3.0, 0.5, 0.5)
    force.addParticleParameterOffset("p2", 1, 1.0, 1.0, 2.0)
    force.addExceptionParameterOffset("p1", 1, 0.5, 0.5, 1.5)
    system.addForce(force)
    context = Context(system, platform)
    context.setPositions([[i, 0, 0] for i in range(4)])
    assert len(context.getParameters()) == 2
    assert context.getParameter("p1") == 0.0
    assert context.getParameter("p2") == 1.0
    context.setParameter("p1", 0.5)
    context.setParameter("p2", 1.5)
    particleCharge = [0.0+3.0*0.5, 1.0+1.0*1.5, -1.0, 0.5]
    particleSigma = [1.0+0.5*0.5, 0.5+1.0*1.5, 2.0, 2.0]
    particleEpsilon = [0.5+0.5*0.5, 0.6+2.0*1.5, 0.7, 0.8]
    qq, sig, eps = np.zeros((4, 4)), np.zeros((4, 4)), np.zeros((4, 4))
    for i in range(4):
        for j in range(i+1, 4):
            qq[i][j] = particleCharge[i]*particleCharge[j]
            sig[i][j] = 0.5*(particleSigma[i]+particleSigma[j])
            eps[i][j] = math.sqrt(particleEpsilon[i]*particleEpsilon[j])
    qq[0][3], sig[0][3], eps[0][3] = 0.0, 1.0, 0.0
    qq[2][3], sig[2][3], eps[2][3] = 0.5+0.5*0.5, 1.0+0.5*0.5, 1.5+1.5*0.5
    qq[0][1], sig[0][1], eps[0][1] = 1.0, 1.5, 1.0
    energy = 0.0
    for i in range(4):
        for j in range(i+1, 4):
            dist = j-i
            x = sig[i][j]/dist
            energy += ONE_4PI_EPS0*qq[i][j]/dist+4.0*eps[i][j]*(x**12-x**6)
    assert_equal_to(energy, context.getState(State.Energy).getPotentialEnergy(), 1e-4)


def _four_particles(method, cutoff=1.0):
    system = System()
    for _ in range(4):
        system.addParticle(1.0)
    system.setDefaultPeriodicBoxVectors([2, 0, 0], [0, 2, 0], [0, 0, 2])
    force = SlicedNonbondedForce(1)
    system.addForce(force)
    force.setNonbondedMethod(method)
    force.setCutoffDistance(cutoff)
    force.addParticle(1.0, 0.5, 1.0)
    force.addParticle(1.0, 0.5, 1.0)
    force.addParticle(-1.0, 0.5, 1.0)
    force.addParticle(-1.0, 0.5, 1.0)
    positions = [[0, 0, 0], [1.5, 0, 0], [0, 0.5, 0.5], [0.2, 1.3, 0]]
    return system, force, positions


def test_ewald_exceptions(platform):
    # testEwaldExceptions (LJPME), :903-941
    system, force, positions = _four_particles(SlicedNonbondedForce.LJPME)
    context = Context(system, platform)
    context.setPositions(positions)
    e1 = context.getState(State.Energy).getPotentialEnergy()
    force.addException(0, 1, 0.2, 0.8, 2.0)
    force.setExceptionsUsePeriodicBoundaryConditions(True)
    context.reinitialize(True)
    e2 = context.getState(State.Energy).getPotentialEnergy()
    r = 0.5
    expectedChange = ONE_4PI_EPS0*(0.2-1.0)/r+4*2.0*((0.8/r)**12-(0.8/r)**6)-4*1.0*((0.5/r)**12-(0.5/r)**6)
    assert_equal_to(expectedChange, e2-e1, 1e-4)


def test_direct_and_reciprocal(platform):
    # testDirectAndReciprocal, :943-985
    system, force, positions = _four_particles(SlicedNonbondedForce.PME)
    force.setReciprocalSpaceForceGroup(1)
    force.addException(0, 2, -2.0, 0.5, 3.0)
    context = Context(system, platform)
    context.setPositions(positions)
    e1 = context.getState(State.Energy).getPotentialEnergy()
    e2 = context.getState(State.Energy, True, 1 << 0).getPotentialEnergy()
    e3 = context.getState(State.Energy, True, 1 << 1).getPotentialEnergy()
    assert_equal_to(e1, e2+e3, 1e-4)
    assert e2 != 0
    assert e3 != 0
    force.setIncludeDirectSpace(False)
    context.reinitialize(True)
    e4 = context.getState(State.Energy).getPotentialEnergy()
    assert_equal_to(e3, e4, 1e-4)


def test_pme_parameter_queries(platform):
    # getPMEParameters / getLJPMEParameters error behaviour, ReferenceOpenMMLabKernels.cpp:314-330
    system, force, positions = _four_particles(SlicedNonbondedForce.PME)
    context = Context(system, platform)
    alpha, nx, ny, nz = force.getPMEParametersInContext(context)
    assert_equal_to(math.sqrt(-math.log(2*5e-4))/1.0, alpha, 1e-12)
    expected = max(int(math.ceil(2*alpha*2.0/(3*5e-4**0.2))), 6)
    assert (nx, ny, nz) == (expected,)*3
    with pytest.raises(mm.SnbError):
        force.getLJPMEParametersInContext(context)
    force.setNonbondedMethod(SlicedNonbondedForce.CutoffPeriodic)
    context.reinitialize(True)
    with pytest.raises(mm.SnbError):
        force.getPMEParametersInContext(context)


def test_box_too_small(platform):
    # ReferenceOpenMMLabKernels.cpp:208-210
    system, force, positions = _four_particles(SlicedNonbondedForce.CutoffPeriodic)
    context = Context(system, platform)
    context.setPositions(positions)
    context.getState(State.Energy)
    context.setPeriodicBoxVectors([1.9, 0, 0], [0, 2, 0], [0, 0, 2])
    with pytest.raises(mm.SnbError) as err:
        context.getState(State.Energy)
    assert err.value.code == mm._abi.ERR_BOX_TOO_SMALL


def test_update_parameters_errors(platform):
    # copyParametersToContext checks, ReferenceOpenMMLabKernels.cpp:264-291
    system, force, positions = _four_particles(SlicedNonbondedForce.CutoffPeriodic)
    force.addException(0, 1, 0.5, 1.0, 1.0)
    force.addException(2, 3, 0.0, 1.0, 0.0)
    context = Context(system, platform)
    force.setExceptionParameters(1, 2, 3, 0.3, 1.0, 0.0)   # an excluded pair becomes a 1-4
    with pytest.raises(mm.SnbError) as err:
        force.updateParametersInContext(context)
    assert err.value.code == mm._abi.ERR_EXCEPTION_SET
    force.setExceptionParameters(1, 2, 3, 0.0, 1.0, 0.0)
    force.addParticle(0, 1, 0)
    with pytest.raises(mm.SnbError) as err:
        force.updateParametersInContext(context)
    assert err.value.code == mm._abi.ERR_PARTICLE_COUNT
