"""Error behaviour at Context creation: the checks of SlicedNonbondedForceImpl::initialize
(openmmapi/src/SlicedNonbondedForceImpl.cpp:36-131), one test per rule with the reference's message, plus the reference's
Python test testParameterClash (python/tests/TestSlicedNonbondedForce.py:52-68).  Host logic: runs on the CPU platforms."""
import pytest

from openmmlab_b200 import Context, SlicedNonbondedForce, System
from openmmlab_b200.force import OpenMMException
from oracle import oracle_kernel  # noqa: F401  (registers the CPU platforms)

from util import CPU_PLATFORMS

F = SlicedNonbondedForce
PLATFORM = CPU_PLATFORMS[-1]


def _pair(method=F.NoCutoff, box=4.0):
    system = System()
    system.setDefaultPeriodicBoxVectors([box, 0, 0], [0, box, 0], [0, 0, box])
    system.addParticle(1.0)
    system.addParticle(1.0)
    force = F(1)
    force.setNonbondedMethod(method)
    force.addParticle(1.5, 1, 0)
    force.addParticle(-1.5, 1, 0)
    system.addForce(force)
    return system, force


def _raises(system, message):
    with pytest.raises(OpenMMException) as err:
        Context(system, PLATFORM)
    assert message in str(err.value)


def test_parameter_clash():
    # testParameterClash: one global parameter as slice scaling AND as parameter offset
    system, force = _pair()
    force.addGlobalParameter("param", 1)
    force.addScalingParameter("param", 0, 0, True, True)
    force.addParticleParameterOffset("param", 0, 1, 1, 0)
    _raises(system, "Cannot use a global parameter for both slice energy scaling and parameter offset")


def test_particle_count_must_match_the_system():
    system, force = _pair()
    force.addParticle(0.0, 1, 0)
    _raises(system, "must have exactly as many particles as the System")


def test_switching_distance_range():
    system, force = _pair(F.CutoffNonPeriodic)
    force.setCutoffDistance(1.0)
    force.setUseSwitchingFunction(True)
    force.setSwitchingDistance(1.0)
    _raises(system, "Switching distance must satisfy 0 <= r_switch < r_cutoff")


@pytest.mark.parametrize("sigma,epsilon,message", [(-1.0, 0.0, "sigma for a particle cannot be negative"),
                                                   (1.0, -1.0, "epsilon for a particle cannot be negative")])
def test_negative_particle_parameters(sigma, epsilon, message):
    system, force = _pair()
    force.setParticleParameters(0, 1.5, sigma, epsilon)
    _raises(system, message)


@pytest.mark.parametrize("sigma,epsilon,message", [(-1.0, 0.0, "sigma for an exception cannot be negative"),
                                                   (1.0, -1.0, "epsilon for an exception cannot be negative")])
def test_negative_exception_parameters(sigma, epsilon, message):
    system, force = _pair()
    force.addException(0, 1, 0.0, sigma, epsilon)
    _raises(system, message)


def test_cutoff_against_half_the_box():
    system, force = _pair(F.CutoffPeriodic, box=4.0)
    force.setCutoffDistance(2.1)
    _raises(system, "The cutoff distance cannot be greater than half the periodic box size")


def test_ewald_needs_a_rectangular_box():
    system, force = _pair(F.Ewald)
    system.setDefaultPeriodicBoxVectors([4, 0, 0], [1, 4, 0], [0, 0, 4])
    _raises(system, "Ewald is not supported with non-rectangular boxes")


def test_offset_indices():
    system, force = _pair()
    force.addGlobalParameter("p", 0.0)
    force.addExceptionParameterOffset("p", 0, 1, 1, 1)      # no exception 0
    _raises(system, "Illegal exception index for an exception parameter offset: 0")
