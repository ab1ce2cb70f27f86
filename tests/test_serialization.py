"""the XML wire format of a SlicedNonbondedForce.

test_serialization is the reference's serialization/tests/TestSerializeSlicedNonbondedForce.cpp:22-180 restated
getter for getter; the fixture test reads a document laid out the way OpenMM's XmlSerializer writes the reference
proxy's node tree (serialization/src/SlicedNonbondedForceProxy.cpp:23-101); the round-trip tests check that a force
that went through XML evaluates to the very same numbers (CPU oracle platforms here, the CUDA path under -m gpu)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import openmmlab_b200 as mm  # noqa: E402
from openmmlab_b200 import systems  # noqa: E402
from openmmlab_b200.serialization import SerializationNode, SlicedNonbondedForceProxy, XmlSerializer  # noqa: E402
from util import PLATFORMS, assert_parity, evaluate  # noqa: E402

F = mm.SlicedNonbondedForce


def reference_test_force():
    # TestSerializeSlicedNonbondedForce.cpp:25-58
    force = F(3)
    force.setForceGroup(3)
    force.setName("custom name")
    force.setNonbondedMethod(F.CutoffPeriodic)
    force.setSwitchingDistance(1.5)
    force.setUseSwitchingFunction(True)
    force.setCutoffDistance(2.0)
    force.setEwaldErrorTolerance(1e-3)
    force.setReactionFieldDielectric(50.0)
    force.setUseDispersionCorrection(False)
    force.setExceptionsUsePeriodicBoundaryConditions(True)
    force.setIncludeDirectSpace(False)
    force.setPMEParameters(0.5, 3, 5, 7)
    force.setLJPMEParameters(0.8, 4, 6, 7)
    force.addParticle(1, 0.1, 0.01)
    force.addParticle(0.5, 0.2, 0.02)
    force.addParticle(-0.5, 0.3, 0.03)
    force.setParticleSubset(0, 1)
    force.setParticleSubset(1, 2)
    force.addException(0, 1, 2, 0.5, 0.1)
    force.addException(1, 2, 0.2, 0.4, 0.2)
    force.addGlobalParameter("scale1", 1.0)
    force.addGlobalParameter("scale2", 2.0)
    force.addParticleParameterOffset("scale1", 2, 1.5, 2.0, 2.5)
    force.addExceptionParameterOffset("scale2", 1, -0.1, -0.2, -0.3)
    force.addGlobalParameter("lambda", 0.5)
    force.addScalingParameter("lambda", 0, 1, True, True)
    force.addScalingParameter("lambda", 1, 1, False, True)
    force.addScalingParameterDerivative("lambda")
    return force


def assert_same_force(force, force2):
    # TestSerializeSlicedNonbondedForce.cpp:68-174 (every ASSERT_EQUAL of the reference's test, plus the ones it leaves out)
    assert force.getNumSubsets() == force2.getNumSubsets()
    assert force.getForceGroup() == force2.getForceGroup()
    assert force.getName() == force2.getName()
    assert force.getNonbondedMethod() == force2.getNonbondedMethod()
    assert force.getSwitchingDistance() == force2.getSwitchingDistance()
    assert force.getUseSwitchingFunction() == force2.getUseSwitchingFunction()
    assert force.getCutoffDistance() == force2.getCutoffDistance()
    assert force.getEwaldErrorTolerance() == force2.getEwaldErrorTolerance()
    assert force.getReactionFieldDielectric() == force2.getReactionFieldDielectric()
    assert force.getUseDispersionCorrection() == force2.getUseDispersionCorrection()
    assert force.getExceptionsUsePeriodicBoundaryConditions() == force2.getExceptionsUsePeriodicBoundaryConditions()
    assert force.getIncludeDirectSpace() == force2.getIncludeDirectSpace()
    assert force.getReciprocalSpaceForceGroup() == force2.getReciprocalSpaceForceGroup()
    assert force.getPMEParameters() == force2.getPMEParameters()
    assert force.getLJPMEParameters() == force2.getLJPMEParameters()
    assert force.getNumParticles() == force2.getNumParticles()
    assert force.getNumExceptions() == force2.getNumExceptions()
    assert force.getNumGlobalParameters() == force2.getNumGlobalParameters()
    assert force.getNumParticleParameterOffsets() == force2.getNumParticleParameterOffsets()
    assert force.getNumExceptionParameterOffsets() == force2.getNumExceptionParameterOffsets()
    for i in range(force.getNumGlobalParameters()):
        assert force.getGlobalParameterName(i) == force2.getGlobalParameterName(i)
        assert force.getGlobalParameterDefaultValue(i) == force2.getGlobalParameterDefaultValue(i)
    for i in range(force.getNumParticleParameterOffsets()):
        assert force.getParticleParameterOffset(i) == force2.getParticleParameterOffset(i)
    for i in range(force.getNumExceptionParameterOffsets()):
        assert force.getExceptionParameterOffset(i) == force2.getExceptionParameterOffset(i)
    for i in range(force.getNumParticles()):
        assert force.getParticleParameters(i) == force2.getParticleParameters(i)
        assert force.getParticleSubset(i) == force2.getParticleSubset(i)
    for i in range(force.getNumExceptions()):
        assert force.getExceptionParameters(i) == force2.getExceptionParameters(i)
    assert force.getNumScalingParameters() == force2.getNumScalingParameters()
    for i in range(force.getNumScalingParameters()):
        assert force.getScalingParameter(i) == force2.getScalingParameter(i)
    assert force.getNumScalingParameterDerivatives() == force2.getNumScalingParameterDerivatives()
    for i in range(force.getNumScalingParameterDerivatives()):
        assert force.getScalingParameterDerivativeName(i) == force2.getScalingParameterDerivativeName(i)


def test_serialization():
    force = reference_test_force()
    text = XmlSerializer.serialize(force, "Force")
    assert_same_force(force, XmlSerializer.deserialize(text))
    # and the document is stable: what was read writes the same bytes again
    assert XmlSerializer.serialize(XmlSerializer.deserialize(text)) == text


# The reference test's force as OpenMM's XmlSerializer lays out the proxy's node tree: attributes in alphabetical order,
# booleans as 0 / 1, shortest doubles without a leading zero, tab indentation, type + openmmVersion on the root.
FIXTURE = """<?xml version="1.0" ?>
<Force alpha=".5" cutoff="2" dispersionCorrection="0" ewaldTolerance=".001" exceptionsUsePeriodic="1" forceGroup="3" includeDirectSpace="0" ljAlpha=".8" ljnx="4" ljny="6" ljnz="7" method="2" name="custom name" numSubsets="3" nx="3" ny="5" nz="7" openmmVersion="8.1" recipForceGroup="-1" rfDielectric="50" switchingDistance="1.5" type="SlicedNonbondedForce" useSwitchingFunction="1" version="1">
	<GlobalParameters>
		<Parameter default="1" name="scale1"/>
		<Parameter default="2" name="scale2"/>
		<Parameter default=".5" name="lambda"/>
	</GlobalParameters>
	<ParticleOffsets>
		<Offset eps="2.5" parameter="scale1" particle="2" q="1.5" sig="2"/>
	</ParticleOffsets>
	<ExceptionOffsets>
		<Offset eps="-.3" exception="1" parameter="scale2" q="-.1" sig="-.2"/>
	</ExceptionOffsets>
	<Particles>
		<Particle eps=".01" q="1" sig=".1"/>
		<Particle eps=".02" q=".5" sig=".2"/>
		<Particle eps=".03" q="-.5" sig=".3"/>
	</Particles>
	<Exceptions>
		<Exception eps=".1" p1="0" p2="1" q="2" sig=".5"/>
		<Exception eps=".2" p1="1" p2="2" q=".2" sig=".4"/>
	</Exceptions>
	<Subsets>
		<Subset index="0" subset="1"/>
		<Subset index="1" subset="2"/>
	</Subsets>
	<scalingParameters>
		<scalingParameter includeCoulomb="1" includeLJ="1" parameter="lambda" subset1="0" subset2="1"/>
		<scalingParameter includeCoulomb="0" includeLJ="1" parameter="lambda" subset1="1" subset2="1"/>
	</scalingParameters>
	<scalingParameterDerivatives>
		<scalingParameterDerivative parameter="lambda"/>
	</scalingParameterDerivatives>
</Force>
"""


def test_fixture_document():
    force = reference_test_force()
    assert_same_force(force, XmlSerializer.deserialize(FIXTURE))
    assert XmlSerializer.serialize(force, "Force") == FIXTURE


def test_optional_properties_and_errors():
    # the properties deserialize() reads with a default (SlicedNonbondedForceProxy.cpp:110-126) may be absent
    node = SerializationNode("Force")
    SlicedNonbondedForceProxy().serialize(reference_test_force(), node)
    for name in ("forceGroup", "name", "useSwitchingFunction", "switchingDistance", "includeDirectSpace", "alpha", "nx", "ny", "nz",
                 "ljAlpha", "ljnx", "ljny", "ljnz", "recipForceGroup"):
        del node.properties[name]
    force = SlicedNonbondedForceProxy().deserialize(node)
    assert force.getForceGroup() == 0 and force.getName() == "SlicedNonbondedForce"
    assert force.getUseSwitchingFunction() is False and force.getSwitchingDistance() == -1.0
    assert force.getIncludeDirectSpace() is True
    assert force.getPMEParameters() == (0.0, 0, 0, 0) and force.getLJPMEParameters() == (0.0, 0, 0, 0)
    assert force.getReciprocalSpaceForceGroup() == -1
    # the mandatory ones may not
    del node.properties["cutoff"]
    with pytest.raises(mm.OpenMMException):
        SlicedNonbondedForceProxy().deserialize(node)
    # version (:104-106)
    with pytest.raises(mm.OpenMMException, match="Unsupported version number"):
        XmlSerializer.deserialize(FIXTURE.replace('version="1">', 'version="2">'))
    with pytest.raises(mm.OpenMMException, match="no serialization proxy"):
        XmlSerializer.deserialize(FIXTURE.replace('type="SlicedNonbondedForce"', 'type="NonbondedForce"'))
    # what the API refuses it refuses on the way in as well: a clash between scaling parameters (SlicedNonbondedForce.cpp)
    with pytest.raises(mm.OpenMMException, match="Clash"):
        XmlSerializer.deserialize(FIXTURE.replace('includeCoulomb="0" includeLJ="1" parameter="lambda" subset1="1" subset2="1"',
                                                  'includeCoulomb="0" includeLJ="1" parameter="lambda" subset1="1" subset2="0"'))


def test_doubles_round_trip_bit_for_bit():
    rng = np.random.default_rng(5)
    values = np.concatenate([rng.normal(size=200), 10.0**rng.uniform(-300, 300, size=200), [0.0, -0.0, 1e-5, 5e-4, 0.1, 1/3, 2.0**-1074]])
    force = F(1)
    for v in values:
        force.addParticle(v, abs(v), 0.25*abs(v))
    copy = XmlSerializer.deserialize(XmlSerializer.serialize(force))
    for i, v in enumerate(values):
        q, s, e = copy.getParticleParameters(i)
        assert np.float64(q).tobytes() == np.float64(v).tobytes() and s == abs(v) and e == 0.25*abs(v)


def _through_xml(w):
    force2 = XmlSerializer.deserialize(XmlSerializer.serialize(w.force))
    assert_same_force(w.force, force2)
    system = mm.System()
    for _ in range(w.num_particles):
        system.addParticle(1.0)
    system.setDefaultPeriodicBoxVectors(*w.system.getDefaultPeriodicBoxVectors())
    system.addForce(force2)
    return systems.Workload(w.name+"_xml", system, force2, w.positions, w.box, w.parameters, w.meta)


@pytest.mark.parametrize("platform", PLATFORMS)
@pytest.mark.parametrize("workload", ["water_pme", "fep_ljpme"])
def test_force_through_xml_evaluates_identically(platform, workload):
    w = systems.water_box() if workload == "water_pme" else systems.fep_ligand()
    if platform == "B200":
        # same parameters bit for bit; deterministic mode accumulates forces and grid charges in fixed point, so two evaluations
        # give the same forces (slice energies are summed with FP64 atomics: order-dependent last bits)
        props = {"DeterministicForces": "true"}
        a = evaluate(platform, w, properties=props)
        b = evaluate(platform, _through_xml(w), properties=props)
        assert np.array_equal(a["forces"], b["forces"])
        assert_parity(a, b, tol=1e-12, what="through XML")
        return
    a = evaluate(platform, w)
    b = evaluate(platform, _through_xml(w))
    # the parameters are bit-identical, so is every result
    assert a["energy"] == b["energy"]
    assert np.array_equal(a["slices"], b["slices"])
    assert a["derivs"] == b["derivs"]
    assert np.array_equal(a["forces"], b["forces"])


def test_python_surface_extras():
    # python/openmmlab.i:221-232 (both constructors), :471-479 (%extend cast / isinstance)
    base = F(1)
    base.setNonbondedMethod(F.PME)
    base.setCutoffDistance(0.9)
    base.addParticle(0.5, 0.3, 0.1)
    base.addParticle(-0.5, 0.3, 0.1)
    base.addException(0, 1, 0.0, 1.0, 0.0)
    base.addGlobalParameter("p", 1.0)
    base.addParticleParameterOffset("p", 0, 0.1, 0.0, 0.0)
    sliced = F(base, 2)                       # SlicedNonbondedForce(force, numSubsets)
    assert sliced.getNumSubsets() == 2 and sliced.getNumSlices() == 3
    assert sliced.getNonbondedMethod() == F.PME and sliced.getCutoffDistance() == 0.9
    assert sliced.getNumParticles() == 2 and sliced.getNumExceptions() == 1
    assert sliced.getParticleParameterOffset(0) == ("p", 0, 0.1, 0.0, 0.0)
    assert F.isinstance(sliced) and not F.isinstance(object())
    assert F.cast(sliced) is sliced
    with pytest.raises(TypeError):
        F.cast(object())
    assert_same_force(sliced, XmlSerializer.deserialize(XmlSerializer.serialize(sliced)))


HAVE_REFERENCE = os.path.isfile("/root/reference/serialization/src/SlicedNonbondedForceProxy.cpp")


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference's sources are not on this machine")
def test_reference_proxy_and_its_own_test_on_the_stand_in_serializer(tmp_path):
    """plugin/Makefile run-serialization-test: the reference's OWN proxy (serialization/src/SlicedNonbondedForceProxy.cpp), its
    registration unit and its serialization test (serialization/tests/TestSerializeSlicedNonbondedForce.cpp), compiled unchanged,
    on the stand-in XmlSerializer of plugin/mock_openmm -- and the cross-language check: what that proxy writes is, byte for byte,
    what serialization.py writes; each side reads the other's documents."""
    import subprocess
    plugin = os.path.join(ROOT, "plugin")
    out = subprocess.run(["make", "-C", plugin, "-B", "run-serialization-test"], capture_output=True, text=True)
    assert out.returncode == 0 and "\nDone\n" in out.stdout, out.stdout[-2000:]+out.stderr[-2000:]
    exe = os.path.join(plugin, "build", "xml_cross")
    # C++ -> Python
    written = str(tmp_path/"from_cpp.xml")
    subprocess.run([exe, "write", written], check=True)
    text = open(written).read()
    assert text == FIXTURE == XmlSerializer.serialize(reference_test_force(), "Force")
    assert_same_force(reference_test_force(), XmlSerializer.deserialize(text))
    # Python -> C++ -> Python, on forces with awkward doubles: the FEP workload's force and random parameters
    rng = np.random.default_rng(11)
    forces = [systems.fep_ligand().force, systems.water_box().force]
    odd = F(2)
    for v in np.concatenate([rng.normal(size=40), 10.0**rng.uniform(-12, 12, size=40), [1e-5, 1e-6, 123456789.125, 1e15, 1e16, 1e17, 0.1, -0.0]]):
        odd.addParticle(v, abs(v)+0.1, 0.5*abs(v))
    odd.setParticleSubset(3, 1)
    odd.addGlobalParameter("x", 1/3)
    odd.addScalingParameter("x", 0, 1, True, False)
    forces.append(odd)
    for k, force in enumerate(forces):
        mine = XmlSerializer.serialize(force)
        src, dst = str(tmp_path/("py%d.xml" % k)), str(tmp_path/("cpp%d.xml" % k))
        open(src, "w").write(mine)
        done = subprocess.run([exe, "copy", src, dst], capture_output=True, text=True)
        assert done.returncode == 0, done.stdout
        theirs = open(dst).read()
        assert_same_force(force, XmlSerializer.deserialize(theirs))
        assert theirs == mine, "the two writers disagree on the bytes of force %d" % k
