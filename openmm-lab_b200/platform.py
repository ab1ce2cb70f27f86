"""Kernel-factory registration for platform "B200".

Mirrors the reference's plugin entry points
(platforms/cuda/src/CudaOpenMMLabKernelFactory.cpp:19-57): ``registerKernelFactories``
registers a factory for the kernel name "CalcSlicedNonbondedForce"; the factory throws for
any other name (:56).  Platform properties follow the reference's CUDA platform
(CudaOpenMMLabKernels.cpp:483,496; platforms/cuda/tests/CudaOpenMMLabTests.h:26-31):
``Precision`` (mixed | double), ``DeterministicForces``, ``DeviceIndex``; plus ``DirectKernel``
(auto | packed | general: include/slicednb.h SNB_FLAG_DIRECT_*).
"""
from .context import Platform
from .force import OpenMMException
from .kernel import B200CalcSlicedNonbondedForceKernel, CalcSlicedNonbondedForceKernel


def _createKernelImpl(name, platform, context):
    if name == CalcSlicedNonbondedForceKernel.Name():
        props = getattr(context, "_properties", platform._properties)
        precision = str(props.get("Precision", "mixed")).lower()
        if precision not in ("mixed", "double"):
            raise OpenMMException("Illegal value for Precision: %s (this platform offers mixed and double)" % precision)
        deterministic = str(props.get("DeterministicForces", "false")).lower() in ("true", "1")
        direct_kernel = str(props.get("DirectKernel", "auto")).lower()
        if direct_kernel not in ("auto", "packed", "general"):
            raise OpenMMException("Illegal value for DirectKernel: %s (auto, packed or general)" % direct_kernel)
        return B200CalcSlicedNonbondedForceKernel(int(props.get("DeviceIndex", 0)), precision, deterministic, direct_kernel)
    raise OpenMMException("Tried to create kernel with illegal kernel name '%s'" % name)


def registerPlatforms():
    if "B200" not in Platform._platforms:
        Platform.registerPlatform(Platform("B200"))


def registerKernelFactories():
    registerPlatforms()
    Platform.getPlatformByName("B200").registerKernelFactory(CalcSlicedNonbondedForceKernel.Name(), _createKernelImpl)
