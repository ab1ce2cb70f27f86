"""The on-disk XML form of a ``SlicedNonbondedForce``: the one wire format adjacent to the path.

Follows the reference's serialization proxy property for property and child for child
(serialization/src/SlicedNonbondedForceProxy.cpp:23-101 ``serialize``, :103-162 ``deserialize``; registered under
the type name "SlicedNonbondedForce", serialization/src/OpenMMLabSerializationProxyRegistration.cpp), written the way
OpenMM's ``XmlSerializer`` writes a ``SerializationNode`` tree (third party, OpenMM 8.1, restated from its published
behaviour: properties become attributes in ``std::map`` = alphabetical order, booleans are stored as the integers
0 / 1, doubles in the shortest form that reads back to the same bits, the root element carries ``type`` and
``openmmVersion``, children are indented with tabs).  So a force serialized by the reference plugin loads into the host
mirror here (``deserialize``) and runs on the B200 kernel, and what ``serialize`` writes is what the reference's
``deserialize`` reads: same names, same defaults for the optional properties (:110-115,126), the same
"Unsupported version number" error for any version but 1 (:104-106).

Host-side only: nothing here touches the CUDA library.
"""
import xml.etree.ElementTree as ET

from .force import OpenMMException, SlicedNonbondedForce

TYPE_NAME = "SlicedNonbondedForce"
VERSION = 1
OPENMM_VERSION = "8.1"


def _fmt_double(x):
    """Shortest decimal string that reads back to the same double, in the style of the ``g_fmt`` OpenMM uses
    (no leading zero before the decimal point; integers without a fraction)."""
    x = float(x)
    if x != x:
        return "nan"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    s = repr(x)
    if s.endswith(".0"):
        s = s[:-2]
    if s.startswith("0."):
        s = s[1:]
    elif s.startswith("-0."):
        s = "-"+s[2:]
    return s


class SerializationNode:
    """The part of OpenMM's ``SerializationNode`` the proxy uses (third party; same method names)."""

    def __init__(self, name=""):
        self.name = name
        self.properties = {}
        self.children = []

    # --- writing
    def setIntProperty(self, name, value):
        self.properties[name] = str(int(value))
        return self

    def setBoolProperty(self, name, value):
        self.properties[name] = "1" if value else "0"
        return self

    def setDoubleProperty(self, name, value):
        self.properties[name] = _fmt_double(value)
        return self

    def setStringProperty(self, name, value):
        self.properties[name] = str(value)
        return self

    def createChildNode(self, name):
        child = SerializationNode(name)
        self.children.append(child)
        return child

    # --- reading
    _missing = object()

    def hasProperty(self, name):
        return name in self.properties

    def _get(self, name, default, convert):
        if name not in self.properties:
            if default is SerializationNode._missing:
                raise OpenMMException("Unknown property '"+name+"' in node '"+self.name+"'")
            return default
        try:
            return convert(self.properties[name])
        except ValueError:
            raise OpenMMException("Illegal value for property '"+name+"' in node '"+self.name+"'")

    def getIntProperty(self, name, default=_missing):
        return self._get(name, default, int)

    def getBoolProperty(self, name, default=_missing):
        return self._get(name, default, lambda s: int(s) != 0)

    def getDoubleProperty(self, name, default=_missing):
        return self._get(name, default, float)

    def getStringProperty(self, name, default=_missing):
        return self._get(name, default, str)

    def getChildren(self):
        return self.children

    def getChildNode(self, name):
        for child in self.children:
            if child.name == name:
                return child
        raise OpenMMException("Unknown child '"+name+"' in node '"+self.name+"'")


class SlicedNonbondedForceProxy:
    """serialization/include/SlicedNonbondedForceProxy.h: ``serialize(object, node)`` / ``deserialize(node)``."""

    typeName = TYPE_NAME

    def serialize(self, force, node):
        # SlicedNonbondedForceProxy.cpp:23-101, in its order
        node.setIntProperty("version", VERSION)
        node.setIntProperty("numSubsets", force.getNumSubsets())
        node.setIntProperty("forceGroup", force.getForceGroup())
        node.setStringProperty("name", force.getName())
        node.setIntProperty("method", force.getNonbondedMethod())
        node.setDoubleProperty("cutoff", force.getCutoffDistance())
        node.setBoolProperty("useSwitchingFunction", force.getUseSwitchingFunction())
        node.setDoubleProperty("switchingDistance", force.getSwitchingDistance())
        node.setDoubleProperty("ewaldTolerance", force.getEwaldErrorTolerance())
        node.setDoubleProperty("rfDielectric", force.getReactionFieldDielectric())
        node.setIntProperty("dispersionCorrection", force.getUseDispersionCorrection())
        node.setIntProperty("exceptionsUsePeriodic", force.getExceptionsUsePeriodicBoundaryConditions())
        node.setBoolProperty("includeDirectSpace", force.getIncludeDirectSpace())
        alpha, nx, ny, nz = force.getPMEParameters()
        node.setDoubleProperty("alpha", alpha).setIntProperty("nx", nx).setIntProperty("ny", ny).setIntProperty("nz", nz)
        alpha, nx, ny, nz = force.getLJPMEParameters()
        node.setDoubleProperty("ljAlpha", alpha).setIntProperty("ljnx", nx).setIntProperty("ljny", ny).setIntProperty("ljnz", nz)
        node.setIntProperty("recipForceGroup", force.getReciprocalSpaceForceGroup())
        params = node.createChildNode("GlobalParameters")
        for i in range(force.getNumGlobalParameters()):
            params.createChildNode("Parameter").setStringProperty("name", force.getGlobalParameterName(i)) \
                .setDoubleProperty("default", force.getGlobalParameterDefaultValue(i))
        offsets = node.createChildNode("ParticleOffsets")
        for i in range(force.getNumParticleParameterOffsets()):
            parameter, particle, dq, dsig, deps = force.getParticleParameterOffset(i)
            offsets.createChildNode("Offset").setStringProperty("parameter", parameter).setIntProperty("particle", particle) \
                .setDoubleProperty("q", dq).setDoubleProperty("sig", dsig).setDoubleProperty("eps", deps)
        offsets = node.createChildNode("ExceptionOffsets")
        for i in range(force.getNumExceptionParameterOffsets()):
            parameter, exception, dq, dsig, deps = force.getExceptionParameterOffset(i)
            offsets.createChildNode("Offset").setStringProperty("parameter", parameter).setIntProperty("exception", exception) \
                .setDoubleProperty("q", dq).setDoubleProperty("sig", dsig).setDoubleProperty("eps", deps)
        particles = node.createChildNode("Particles")
        for i in range(force.getNumParticles()):
            q, sig, eps = force.getParticleParameters(i)
            particles.createChildNode("Particle").setDoubleProperty("q", q).setDoubleProperty("sig", sig).setDoubleProperty("eps", eps)
        exceptions = node.createChildNode("Exceptions")
        for i in range(force.getNumExceptions()):
            p1, p2, q, sig, eps = force.getExceptionParameters(i)
            exceptions.createChildNode("Exception").setIntProperty("p1", p1).setIntProperty("p2", p2) \
                .setDoubleProperty("q", q).setDoubleProperty("sig", sig).setDoubleProperty("eps", eps)
        subsets = node.createChildNode("Subsets")
        for i in range(force.getNumParticles()):
            subset = force.getParticleSubset(i)
            if subset != 0:             # only the non-default ones are stored (:86-90)
                subsets.createChildNode("Subset").setIntProperty("index", i).setIntProperty("subset", subset)
        scaling = node.createChildNode("scalingParameters")
        for i in range(force.getNumScalingParameters()):
            parameter, s1, s2, coulomb, lj = force.getScalingParameter(i)
            scaling.createChildNode("scalingParameter").setStringProperty("parameter", parameter).setIntProperty("subset1", s1) \
                .setIntProperty("subset2", s2).setBoolProperty("includeCoulomb", coulomb).setBoolProperty("includeLJ", lj)
        derivatives = node.createChildNode("scalingParameterDerivatives")
        for i in range(force.getNumScalingParameterDerivatives()):
            derivatives.createChildNode("scalingParameterDerivative").setStringProperty("parameter", force.getScalingParameterDerivativeName(i))

    def deserialize(self, node):
        # SlicedNonbondedForceProxy.cpp:103-162, in its order (offsets before particles: they are only checked at Context creation)
        if node.getIntProperty("version") != VERSION:
            raise OpenMMException("Unsupported version number")
        force = SlicedNonbondedForce(node.getIntProperty("numSubsets"))
        force.setForceGroup(node.getIntProperty("forceGroup", 0))
        force.setName(node.getStringProperty("name", force.getName()))
        force.setNonbondedMethod(node.getIntProperty("method"))
        force.setCutoffDistance(node.getDoubleProperty("cutoff"))
        force.setUseSwitchingFunction(node.getBoolProperty("useSwitchingFunction", False))
        force.setSwitchingDistance(node.getDoubleProperty("switchingDistance", -1.0))
        force.setEwaldErrorTolerance(node.getDoubleProperty("ewaldTolerance"))
        force.setReactionFieldDielectric(node.getDoubleProperty("rfDielectric"))
        force.setUseDispersionCorrection(node.getIntProperty("dispersionCorrection"))
        if node.hasProperty("includeDirectSpace"):
            force.setIncludeDirectSpace(node.getBoolProperty("includeDirectSpace"))
        force.setPMEParameters(node.getDoubleProperty("alpha", 0.0), node.getIntProperty("nx", 0),
                               node.getIntProperty("ny", 0), node.getIntProperty("nz", 0))
        force.setLJPMEParameters(node.getDoubleProperty("ljAlpha", 0.0), node.getIntProperty("ljnx", 0),
                                 node.getIntProperty("ljny", 0), node.getIntProperty("ljnz", 0))
        force.setReciprocalSpaceForceGroup(node.getIntProperty("recipForceGroup", -1))
        for p in node.getChildNode("GlobalParameters").getChildren():
            force.addGlobalParameter(p.getStringProperty("name"), p.getDoubleProperty("default"))
        for o in node.getChildNode("ParticleOffsets").getChildren():
            force.addParticleParameterOffset(o.getStringProperty("parameter"), o.getIntProperty("particle"),
                                             o.getDoubleProperty("q"), o.getDoubleProperty("sig"), o.getDoubleProperty("eps"))
        for o in node.getChildNode("ExceptionOffsets").getChildren():
            force.addExceptionParameterOffset(o.getStringProperty("parameter"), o.getIntProperty("exception"),
                                              o.getDoubleProperty("q"), o.getDoubleProperty("sig"), o.getDoubleProperty("eps"))
        force.setExceptionsUsePeriodicBoundaryConditions(node.getIntProperty("exceptionsUsePeriodic"))
        for p in node.getChildNode("Particles").getChildren():
            force.addParticle(p.getDoubleProperty("q"), p.getDoubleProperty("sig"), p.getDoubleProperty("eps"))
        for e in node.getChildNode("Exceptions").getChildren():
            force.addException(e.getIntProperty("p1"), e.getIntProperty("p2"), e.getDoubleProperty("q"),
                               e.getDoubleProperty("sig"), e.getDoubleProperty("eps"))
        for s in node.getChildNode("Subsets").getChildren():
            force.setParticleSubset(s.getIntProperty("index"), s.getIntProperty("subset"))
        for p in node.getChildNode("scalingParameters").getChildren():
            force.addScalingParameter(p.getStringProperty("parameter"), p.getIntProperty("subset1"), p.getIntProperty("subset2"),
                                      p.getBoolProperty("includeCoulomb"), p.getBoolProperty("includeLJ"))
        for p in node.getChildNode("scalingParameterDerivatives").getChildren():
            force.addScalingParameterDerivative(p.getStringProperty("parameter"))
        return force


def _escape(s):
    return s.replace("&", "&amp;").replace("<", "&lt;").replace(">", "&gt;").replace('"', "&quot;")


def _write(node, depth, out, extra=()):
    attrs = sorted(list(node.properties.items())+list(extra))
    out.append("\t"*depth+"<"+node.name+"".join(' %s="%s"' % (k, _escape(v)) for k, v in attrs))
    if node.children:
        out.append(">\n")
        for child in node.children:
            _write(child, depth+1, out)
        out.append("\t"*depth+"</"+node.name+">\n")
    else:
        out.append("/>\n")


def _read(element):
    node = SerializationNode(element.tag)
    node.properties = dict(element.attrib)
    node.children = [_read(child) for child in element]
    return node


class XmlSerializer:
    """``XmlSerializer::serialize<SlicedNonbondedForce>(&force, "Force", stream)`` /
    ``XmlSerializer::deserialize<SlicedNonbondedForce>(stream)`` as the reference's test uses them
    (serialization/tests/TestSerializeSlicedNonbondedForce.cpp:62-64), on strings."""

    @staticmethod
    def serialize(force, rootName="Force"):
        node = SerializationNode(rootName)
        SlicedNonbondedForceProxy().serialize(force, node)
        out = ['<?xml version="1.0" ?>\n']
        _write(node, 0, out, extra=(("openmmVersion", OPENMM_VERSION), ("type", TYPE_NAME)))
        return "".join(out)

    @staticmethod
    def deserialize(text):
        try:
            root = ET.fromstring(text)
        except ET.ParseError as e:
            raise OpenMMException("Error parsing XML: %s" % e)
        kind = root.attrib.get("type")
        if kind is None:
            raise OpenMMException("Unknown property 'type' in node '"+root.tag+"'")
        if kind != TYPE_NAME:
            raise OpenMMException("There is no serialization proxy registered for type "+kind)
        node = _read(root)
        node.properties.pop("type", None)
        node.properties.pop("openmmVersion", None)
        return SlicedNonbondedForceProxy().deserialize(node)
