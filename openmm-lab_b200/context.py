"""Just enough of OpenMM's System / Platform / Context / State to drive the kernel
interface the way the reference's tests do (tests/TestSlicedNonbondedForce.h).

OpenMM itself is third party and absent here; these classes carry no physics.  The
one piece of reference logic restated is the force-group routing of
SlicedNonbondedForceImpl::calcForcesAndEnergy (openmmapi/src/SlicedNonbondedForceImpl.cpp:135-142).
"""
import numpy as np

from .force import OpenMMException, validate_for_context


class System:
    def __init__(self):
        self._masses = []
        self._forces = []
        self._box = np.diag([2.0, 2.0, 2.0])

    def addParticle(self, mass):
        self._masses.append(float(mass))
        return len(self._masses)-1

    def getNumParticles(self):
        return len(self._masses)

    def setNumParticles(self, n):
        self._masses = [1.0]*int(n)

    def setDefaultPeriodicBoxVectors(self, a, b, c):
        self._box = np.array([a, b, c], dtype=np.float64)

    def getDefaultPeriodicBoxVectors(self):
        return self._box.copy()

    def addForce(self, force):
        self._forces.append(force)
        return len(self._forces)-1

    def getForces(self):
        return list(self._forces)

    def usesPeriodicBoundaryConditions(self):
        return any(f.usesPeriodicBoundaryConditions() for f in self._forces)


class Platform:
    """Named registry of kernel factories (OpenMM Platform::registerKernelFactory)."""
    _platforms = {}

    def __init__(self, name):
        self._name = name
        self._factories = {}
        self._properties = {}

    def getName(self):
        return self._name

    def registerKernelFactory(self, kernelName, factory):
        self._factories[kernelName] = factory

    def createKernel(self, kernelName, context):
        if kernelName not in self._factories:
            raise OpenMMException("Called createKernel() on a Platform which does not support the requested kernel")
        return self._factories[kernelName](kernelName, self, context)

    @classmethod
    def registerPlatform(cls, platform):
        cls._platforms[platform.getName()] = platform

    @classmethod
    def getPlatformByName(cls, name):
        if name not in cls._platforms:
            raise OpenMMException("There is no registered Platform called \"%s\"" % name)
        return cls._platforms[name]


class State:
    Positions, Velocities, Forces, Energy, Parameters, ParameterDerivatives = 1, 2, 4, 8, 16, 32

    def __init__(self, energy, forces, derivatives):
        self._energy, self._forces, self._derivatives = energy, forces, derivatives

    def getPotentialEnergy(self):
        return self._energy

    def getForces(self):
        return self._forces

    def getEnergyParameterDerivatives(self):
        return self._derivatives


class Context:
    def __init__(self, system, platform, properties=None):
        self._system = system
        self._platform = Platform.getPlatformByName(platform) if isinstance(platform, str) else platform
        # platform properties of this Context (OpenMM: Context(system, integrator, platform, properties))
        self._properties = dict(self._platform._properties)
        self._properties.update(properties or {})
        self._positions = np.zeros((system.getNumParticles(), 3))
        self._parameters = {}
        self._initialize()

    def _initialize(self):
        self._box = self._system.getDefaultPeriodicBoxVectors()
        self._kernels = []
        from .kernel import CalcSlicedNonbondedForceKernel
        for force in self._system.getForces():
            for i in range(force.getNumGlobalParameters()):
                self._parameters.setdefault(force.getGlobalParameterName(i), force.getGlobalParameterDefaultValue(i))
        for force in self._system.getForces():
            # SlicedNonbondedForceImpl::initialize (SlicedNonbondedForceImpl.cpp:33-132)
            kernel = self._platform.createKernel(CalcSlicedNonbondedForceKernel.Name(), self)
            validate_for_context(force, self._system)
            kernel.initialize(self._system, force)
            self._kernels.append((force, kernel))

    def _kernelFor(self, force):
        for f, k in self._kernels:
            if f is force:
                return k
        raise OpenMMException("getImplInContext: This Force is not present in the Context")

    def getSystem(self):
        return self._system

    def getPlatform(self):
        return self._platform

    def setPositions(self, positions):
        positions = np.asarray(positions, dtype=np.float64).reshape(-1, 3)
        if positions.shape[0] != self._system.getNumParticles():
            raise OpenMMException("Called setPositions() on a Context with the wrong number of positions")
        self._positions = positions.copy()

    def setPeriodicBoxVectors(self, a, b, c):
        self._box = np.array([a, b, c], dtype=np.float64)

    def getParameters(self):
        return dict(self._parameters)

    def getParameter(self, name):
        if name not in self._parameters:
            raise OpenMMException("Called getParameter() with invalid parameter name")
        return self._parameters[name]

    def setParameter(self, name, value):
        if name not in self._parameters:
            raise OpenMMException("Called setParameter() with invalid parameter name")
        self._parameters[name] = float(value)

    def reinitialize(self, preserveState=False):
        positions, params, box = self._positions, dict(self._parameters), self._box
        for _, k in self._kernels:
            k.destroy()
        if not preserveState:
            self._parameters = {}
            self._positions = np.zeros((self._system.getNumParticles(), 3))
        self._initialize()
        if preserveState:
            self._positions = positions
            self._box = box
            for name, value in params.items():
                if name in self._parameters:
                    self._parameters[name] = value

    def getState(self, types=State.Energy | State.Forces, enforcePeriodicBox=False, groups=0xFFFFFFFF):
        includeForces = bool(types & State.Forces)
        includeEnergy = bool(types & State.Energy)
        self._forces = np.zeros((self._system.getNumParticles(), 3))
        self._energyParameterDerivatives = {}
        energy = 0.0
        for force, kernel in self._kernels:
            # SlicedNonbondedForceImpl::calcForcesAndEnergy (SlicedNonbondedForceImpl.cpp:135-142)
            includeDirect = force.getIncludeDirectSpace() and (groups & (1 << force.getForceGroup())) != 0
            reciprocalGroup = force.getReciprocalSpaceForceGroup()
            if reciprocalGroup < 0:
                reciprocalGroup = force.getForceGroup()
            includeReciprocal = (groups & (1 << reciprocalGroup)) != 0
            energy += kernel.execute(self, includeForces, includeEnergy, includeDirect, includeReciprocal)
        return State(energy if includeEnergy else 0.0, self._forces.copy(), dict(self._energyParameterDerivatives))
