"""Host-side mirror of the reference's user-facing force object.

``SlicedNonbondedForce`` keeps the method names, argument meaning and error
behaviour of the reference class
(openmmapi/include/SlicedNonbondedForce.h:26-71, openmmapi/src/SlicedNonbondedForce.cpp)
and of the part of OpenMM's ``NonbondedForce`` it inherits, so that tests here read
like tests/TestSlicedNonbondedForce.h.  It is a plain container: it owns no GPU
state.  ``PackedDesc`` (_abi.py) flattens it into the C-ABI descriptor.
"""
import math

from . import _abi
from ._abi import SnbError


class OpenMMException(SnbError):
    def __init__(self, message, code=_abi.ERR_INVALID):
        super().__init__(code, message)


class _ScalingParameterInfo:
    # openmmapi/include/SlicedNonbondedForce.h:77-97
    def __init__(self, global_param_index, subset1, subset2, include_coulomb, include_lj):
        if not (include_coulomb or include_lj):
            raise OpenMMException("Keywords 'includeCoulomb' and 'includeLJ' cannot be both false")
        self.globalParamIndex = global_param_index
        self.subset1 = subset1
        self.subset2 = subset2
        self.includeCoulomb = bool(include_coulomb)
        self.includeLJ = bool(include_lj)

    def getSlice(self):
        i, j = self.subset1, self.subset2
        return i*(i+1)//2+j if i > j else j*(j+1)//2+i

    def clashesWith(self, info):
        return self.getSlice() == info.getSlice() and (
            (self.includeCoulomb and info.includeCoulomb) or (self.includeLJ and info.includeLJ))


class SlicedNonbondedForce:
    NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME, LJPME = range(6)

    def __init__(self, numSubsets, template=None):
        # both constructors of the reference (openmmapi/include/SlicedNonbondedForce.h:29-38, python/openmmlab.i:221-232):
        # SlicedNonbondedForce(numSubsets) and SlicedNonbondedForce(force, numSubsets)
        if not isinstance(numSubsets, (int, float)) and template is not None:
            numSubsets, template = template, numSubsets
        self._numSubsets = int(numSubsets)
        self._method = self.NoCutoff
        self._cutoff = 1.0
        self._useSwitch = False
        self._switchingDistance = -1.0
        self._rfDielectric = 78.3
        self._ewaldErrorTol = 5e-4
        self._alpha, self._nx, self._ny, self._nz = 0.0, 0, 0, 0
        self._dalpha, self._dnx, self._dny, self._dnz = 0.0, 0, 0, 0
        self._useDispersionCorrection = True
        self._exceptionsUsePeriodic = False
        self._recipForceGroup = -1
        self._includeDirectSpace = True
        self._forceGroup = 0
        self._name = "SlicedNonbondedForce"     # OpenMM Force::getName default: the class name
        self._particles = []            # [charge, sigma, epsilon]
        self._exceptions = []           # [p1, p2, chargeProd, sigma, epsilon]
        self._exceptionMap = {}
        self._globalParameters = []     # [name, default]
        self._particleOffsets = []      # [paramIndex, particle, dq, dsigma, deps]
        self._exceptionOffsets = []
        self._subsets = {}
        self._scalingParameters = []
        self._scalingParameterDerivatives = []
        self._useCudaFFT = False
        if template is not None:
            self._copyFrom(template)

    # SlicedNonbondedForce(const NonbondedForce&, int): SlicedNonbondedForce.cpp:35-82
    def _copyFrom(self, f):
        self.setForceGroup(f.getForceGroup())
        self.setNonbondedMethod(f.getNonbondedMethod())
        self.setCutoffDistance(f.getCutoffDistance())
        self.setUseSwitchingFunction(f.getUseSwitchingFunction())
        self.setSwitchingDistance(f.getSwitchingDistance())
        self.setEwaldErrorTolerance(f.getEwaldErrorTolerance())
        self.setReactionFieldDielectric(f.getReactionFieldDielectric())
        self.setUseDispersionCorrection(f.getUseDispersionCorrection())
        self.setIncludeDirectSpace(f.getIncludeDirectSpace())
        self.setPMEParameters(*f.getPMEParameters())
        self.setLJPMEParameters(*f.getLJPMEParameters())
        self.setReciprocalSpaceForceGroup(f.getReciprocalSpaceForceGroup())
        for i in range(f.getNumParticles()):
            self.addParticle(*f.getParticleParameters(i))
        for i in range(f.getNumExceptions()):
            self.addException(*f.getExceptionParameters(i))
        self.setExceptionsUsePeriodicBoundaryConditions(f.getExceptionsUsePeriodicBoundaryConditions())
        for i in range(f.getNumGlobalParameters()):
            self.addGlobalParameter(f.getGlobalParameterName(i), f.getGlobalParameterDefaultValue(i))
        for i in range(f.getNumParticleParameterOffsets()):
            self.addParticleParameterOffset(*f.getParticleParameterOffset(i))
        for i in range(f.getNumExceptionParameterOffsets()):
            self.addExceptionParameterOffset(*f.getExceptionParameterOffset(i))

    # ---- OpenMM NonbondedForce surface (third party; names as used by the reference tests) ----
    def getForceGroup(self): return self._forceGroup
    def setForceGroup(self, g): self._forceGroup = int(g)
    def getName(self): return self._name
    def setName(self, name): self._name = str(name)
    def getNonbondedMethod(self): return self._method
    def setNonbondedMethod(self, m):
        if m not in range(6):
            raise OpenMMException("NonbondedForce: Illegal value for nonbonded method")
        self._method = int(m)
    def getCutoffDistance(self): return self._cutoff
    def setCutoffDistance(self, d): self._cutoff = float(d)
    def getUseSwitchingFunction(self): return self._useSwitch
    def setUseSwitchingFunction(self, u): self._useSwitch = bool(u)
    def getSwitchingDistance(self): return self._switchingDistance
    def setSwitchingDistance(self, d): self._switchingDistance = float(d)
    def getReactionFieldDielectric(self): return self._rfDielectric
    def setReactionFieldDielectric(self, d): self._rfDielectric = float(d)
    def getEwaldErrorTolerance(self): return self._ewaldErrorTol
    def setEwaldErrorTolerance(self, t): self._ewaldErrorTol = float(t)
    def getPMEParameters(self): return self._alpha, self._nx, self._ny, self._nz
    def setPMEParameters(self, alpha, nx, ny, nz):
        self._alpha, self._nx, self._ny, self._nz = float(alpha), int(nx), int(ny), int(nz)
    def getLJPMEParameters(self): return self._dalpha, self._dnx, self._dny, self._dnz
    def setLJPMEParameters(self, alpha, nx, ny, nz):
        self._dalpha, self._dnx, self._dny, self._dnz = float(alpha), int(nx), int(ny), int(nz)
    def getUseDispersionCorrection(self): return self._useDispersionCorrection
    def setUseDispersionCorrection(self, u): self._useDispersionCorrection = bool(u)
    def getExceptionsUsePeriodicBoundaryConditions(self): return self._exceptionsUsePeriodic
    def setExceptionsUsePeriodicBoundaryConditions(self, p): self._exceptionsUsePeriodic = bool(p)
    def getReciprocalSpaceForceGroup(self): return self._recipForceGroup
    def setReciprocalSpaceForceGroup(self, g): self._recipForceGroup = int(g)
    def getIncludeDirectSpace(self): return self._includeDirectSpace
    def setIncludeDirectSpace(self, i): self._includeDirectSpace = bool(i)
    def usesPeriodicBoundaryConditions(self):
        return self._method in (self.CutoffPeriodic, self.Ewald, self.PME, self.LJPME)

    def getNumParticles(self): return len(self._particles)
    def addParticle(self, charge, sigma, epsilon):
        self._particles.append([float(charge), float(sigma), float(epsilon)])
        return len(self._particles)-1
    def getParticleParameters(self, index):
        self._check("Index", index, len(self._particles))
        return tuple(self._particles[index])
    def setParticleParameters(self, index, charge, sigma, epsilon):
        self._check("Index", index, len(self._particles))
        self._particles[index] = [float(charge), float(sigma), float(epsilon)]

    def getNumExceptions(self): return len(self._exceptions)
    def addException(self, particle1, particle2, chargeProd, sigma, epsilon, replace=False):
        key = (min(particle1, particle2), max(particle1, particle2))
        if key in self._exceptionMap:
            if not replace:
                raise OpenMMException("NonbondedForce: There is already an exception for particles %d and %d" % (particle1, particle2))
            self._exceptions[self._exceptionMap[key]] = [particle1, particle2, float(chargeProd), float(sigma), float(epsilon)]
            return self._exceptionMap[key]
        self._exceptions.append([int(particle1), int(particle2), float(chargeProd), float(sigma), float(epsilon)])
        self._exceptionMap[key] = len(self._exceptions)-1
        return len(self._exceptions)-1
    def getExceptionParameters(self, index):
        self._check("Index", index, len(self._exceptions))
        return tuple(self._exceptions[index])
    def setExceptionParameters(self, index, particle1, particle2, chargeProd, sigma, epsilon):
        self._check("Index", index, len(self._exceptions))
        self._exceptions[index] = [int(particle1), int(particle2), float(chargeProd), float(sigma), float(epsilon)]

    def createExceptionsFromBonds(self, bonds, coulomb14Scale, lj14Scale):
        """OpenMM NonbondedForce::createExceptionsFromBonds (third party): 1-2 and 1-3 pairs are
        excluded, 1-4 pairs get scaled exceptions built from the particle parameters."""
        n = self.getNumParticles()
        bonded12 = [set() for _ in range(n)]
        for a, b in bonds:
            bonded12[a].add(b)
            bonded12[b].add(a)
        for i in range(n):
            within2 = set()
            for j in bonded12[i]:
                within2.add(j)
                within2.update(bonded12[j])
            within2.discard(i)
            within3 = set(within2)
            for j in within2:
                within3.update(bonded12[j])
            within3.discard(i)
            for j in sorted(within3):
                if j >= i:
                    continue
                if j in within2:
                    self.addException(j, i, 0.0, 1.0, 0.0)
                else:
                    q1, s1, e1 = self._particles[j]
                    q2, s2, e2 = self._particles[i]
                    self.addException(j, i, coulomb14Scale*q1*q2, 0.5*(s1+s2), lj14Scale*math.sqrt(e1*e2))

    def getNumGlobalParameters(self): return len(self._globalParameters)
    def addGlobalParameter(self, name, defaultValue):
        self._globalParameters.append([str(name), float(defaultValue)])
        return len(self._globalParameters)-1
    def getGlobalParameterName(self, index): return self._globalParameters[index][0]
    def getGlobalParameterDefaultValue(self, index): return self._globalParameters[index][1]
    def setGlobalParameterDefaultValue(self, index, value): self._globalParameters[index][1] = float(value)

    def getNumParticleParameterOffsets(self): return len(self._particleOffsets)
    def addParticleParameterOffset(self, parameter, particleIndex, chargeScale, sigmaScale, epsilonScale):
        self._particleOffsets.append([self._getGlobalParameterIndex(parameter), int(particleIndex), float(chargeScale), float(sigmaScale), float(epsilonScale)])
        return len(self._particleOffsets)-1
    def getParticleParameterOffset(self, index):
        o = self._particleOffsets[index]
        return (self._globalParameters[o[0]][0], o[1], o[2], o[3], o[4])
    def getNumExceptionParameterOffsets(self): return len(self._exceptionOffsets)
    def addExceptionParameterOffset(self, parameter, exceptionIndex, chargeProdScale, sigmaScale, epsilonScale):
        self._exceptionOffsets.append([self._getGlobalParameterIndex(parameter), int(exceptionIndex), float(chargeProdScale), float(sigmaScale), float(epsilonScale)])
        return len(self._exceptionOffsets)-1
    def getExceptionParameterOffset(self, index):
        o = self._exceptionOffsets[index]
        return (self._globalParameters[o[0]][0], o[1], o[2], o[3], o[4])

    # ---- SlicedNonbondedForce surface (openmmapi/include/SlicedNonbondedForce.h:26-71) ----
    def getNonbondedMethodName(self):
        return ["NoCutoff", "CutoffNonPeriodic", "CutoffPeriodic", "Ewald", "PME", "LJPME"][self._method]
    def getNumSubsets(self): return self._numSubsets
    def getNumSlices(self): return self._numSubsets*(self._numSubsets+1)//2
    def getNumScalingParameters(self): return len(self._scalingParameters)
    def getNumScalingParameterDerivatives(self): return len(self._scalingParameterDerivatives)

    @staticmethod
    def _check(name, value, number):
        # ASSERT_VALID, SlicedNonbondedForce.cpp:23-29
        if value < 0 or value >= number:
            raise OpenMMException("%s out of range" % name)

    def setParticleSubset(self, index, subset):
        self._check("Index", index, self.getNumParticles())
        self._check("Subset", subset, self._numSubsets)
        self._subsets[index] = int(subset)

    def getParticleSubset(self, index):
        self._check("Index", index, self.getNumParticles())
        return self._subsets.get(index, 0)

    def _getGlobalParameterIndex(self, parameter):
        for i, (name, _) in enumerate(self._globalParameters):
            if name == parameter:
                return i
        raise OpenMMException("There is no global parameter called '"+parameter+"'")

    def addScalingParameter(self, parameter, subset1, subset2, includeCoulomb, includeLJ):
        self._check("Subset", subset1, self._numSubsets)
        self._check("Subset", subset2, self._numSubsets)
        info = _ScalingParameterInfo(self._getGlobalParameterIndex(parameter), subset1, subset2, includeCoulomb, includeLJ)
        for param in self._scalingParameters:
            if param.clashesWith(info):
                raise OpenMMException("Clash detected between scaling parameters")
        self._scalingParameters.append(info)
        return len(self._scalingParameters)-1

    def getScalingParameter(self, index):
        self._check("Index", index, len(self._scalingParameters))
        info = self._scalingParameters[index]
        return (self._globalParameters[info.globalParamIndex][0], info.subset1, info.subset2, info.includeCoulomb, info.includeLJ)

    def setScalingParameter(self, index, parameter, subset1, subset2, includeCoulomb, includeLJ):
        self._check("Index", index, len(self._scalingParameters))
        self._check("Subset", subset1, self._numSubsets)
        self._check("Subset", subset2, self._numSubsets)
        info = _ScalingParameterInfo(self._getGlobalParameterIndex(parameter), subset1, subset2, includeCoulomb, includeLJ)
        old = self._scalingParameters[index]
        if not old.clashesWith(info):
            for param in self._scalingParameters:
                if param.clashesWith(info):
                    raise OpenMMException("A scaling parameter has already been defined for this slice & contribution(s)")
        self._scalingParameters[index] = info

    def _getScalingParameterIndex(self, parameter):
        for i, info in enumerate(self._scalingParameters):
            if self._globalParameters[info.globalParamIndex][0] == parameter:
                return i
        raise OpenMMException("There is no scaling parameter called '"+parameter+"'")

    def addScalingParameterDerivative(self, parameter):
        index = self._getScalingParameterIndex(parameter)
        if index in self._scalingParameterDerivatives:
            raise OpenMMException("This scaling parameter derivative has already been requested")
        self._scalingParameterDerivatives.append(index)
        return len(self._scalingParameterDerivatives)-1

    def getScalingParameterDerivativeName(self, index):
        self._check("Index", index, len(self._scalingParameterDerivatives))
        info = self._scalingParameters[self._scalingParameterDerivatives[index]]
        return self._globalParameters[info.globalParamIndex][0]

    def setScalingParameterDerivative(self, index, parameter):
        self._check("Index", index, len(self._scalingParameterDerivatives))
        sp = self._getScalingParameterIndex(parameter)
        if self._scalingParameterDerivatives[index] != sp:
            if sp in self._scalingParameterDerivatives:
                raise OpenMMException("This scaling parameter derivative has already been requested")
            self._scalingParameterDerivatives[index] = sp

    # python/openmmlab.i:471-479 (%extend): casting a Force of a System back to its class
    @staticmethod
    def cast(force):
        if not isinstance(force, SlicedNonbondedForce):
            raise TypeError("not a SlicedNonbondedForce")
        return force

    @staticmethod
    def isinstance(force):
        return isinstance(force, SlicedNonbondedForce)

    def getUseCudaFFT(self): return self._useCudaFFT
    def setUseCuFFT(self, use): self._useCudaFFT = bool(use)

    # Context-facing calls; `context` is a context.Context (stands in for OpenMM's)
    def getPMEParametersInContext(self, context):
        return context._kernelFor(self).getPMEParameters()

    def getLJPMEParametersInContext(self, context):
        return context._kernelFor(self).getLJPMEParameters()

    def updateParametersInContext(self, context):
        context._kernelFor(self).copyParametersToContext(context, self)


def validate_for_context(force, system):
    """Checks of SlicedNonbondedForceImpl::initialize (openmmapi/src/SlicedNonbondedForceImpl.cpp:38-131),
    same order, same messages."""
    if system.getNumParticles() != force.getNumParticles():
        raise OpenMMException("SlicedNonbondedForce must have exactly as many particles as the System it belongs to.")
    if force.getUseSwitchingFunction():
        if force.getSwitchingDistance() < 0 or force.getSwitchingDistance() >= force.getCutoffDistance():
            raise OpenMMException("SlicedNonbondedForce: Switching distance must satisfy 0 <= r_switch < r_cutoff")
    n = force.getNumParticles()
    for i in range(n):
        _, sigma, epsilon = force.getParticleParameters(i)
        if sigma < 0:
            raise OpenMMException("SlicedNonbondedForce: sigma for a particle cannot be negative")
        if epsilon < 0:
            raise OpenMMException("SlicedNonbondedForce: epsilon for a particle cannot be negative")
    seen = set()
    for i in range(force.getNumExceptions()):
        p1, p2, _, sigma, epsilon = force.getExceptionParameters(i)
        for p in (p1, p2):
            if p < 0 or p >= n:
                raise OpenMMException("SlicedNonbondedForce: Illegal particle index for an exception: %d" % p)
        key = (min(p1, p2), max(p1, p2))
        if key in seen:
            raise OpenMMException("SlicedNonbondedForce: Multiple exceptions are specified for particles %d and %d" % (p1, p2))
        seen.add(key)
        if sigma < 0:
            raise OpenMMException("SlicedNonbondedForce: sigma for an exception cannot be negative")
        if epsilon < 0:
            raise OpenMMException("SlicedNonbondedForce: epsilon for an exception cannot be negative")
    for i in range(force.getNumParticleParameterOffsets()):
        idx = force.getParticleParameterOffset(i)[1]
        if not 0 <= idx < n:
            raise OpenMMException("SlicedNonbondedForce: Illegal particle index for a particle parameter offset: %d" % idx)
    for i in range(force.getNumExceptionParameterOffsets()):
        idx = force.getExceptionParameterOffset(i)[1]
        if not 0 <= idx < force.getNumExceptions():
            raise OpenMMException("SlicedNonbondedForce: Illegal exception index for an exception parameter offset: %d" % idx)
    if force.usesPeriodicBoundaryConditions():
        box = system.getDefaultPeriodicBoxVectors()
        cutoff = force.getCutoffDistance()
        if cutoff > 0.5*box[0][0] or cutoff > 0.5*box[1][1] or cutoff > 0.5*box[2][2]:
            raise OpenMMException("SlicedNonbondedForce: The cutoff distance cannot be greater than half the periodic box size.")
        if force.getNonbondedMethod() == force.Ewald and (box[1][0] != 0.0 or box[2][0] != 0.0 or box[2][1] != 0.0):
            raise OpenMMException("SlicedNonbondedForce: Ewald is not supported with non-rectangular boxes.  Use PME instead.")
    offset_params = {force.getParticleParameterOffset(i)[0] for i in range(force.getNumParticleParameterOffsets())}
    offset_params |= {force.getExceptionParameterOffset(i)[0] for i in range(force.getNumExceptionParameterOffsets())}
    for i in range(force.getNumScalingParameters()):
        if force.getScalingParameter(i)[0] in offset_params:
            raise OpenMMException("SlicedNonbondedForce: Cannot use a global parameter for both slice energy scaling and parameter offset.")
