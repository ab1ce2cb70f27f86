"""Kernel-factory registration for platform "B200".

Mirrors the reference's plugin entry points
(platforms/cuda/src/CudaOpenMMLabKernelFactory.cpp:19-57): ``registerKernelFactories``
registers a factory for the kernel name "CalcSlicedNonbondedForce"; the factory throws for
any other name (:56).
"""
from .context import Platform
from .force import OpenMMException
from .kernel import B200CalcSlicedNonbondedForceKernel, CalcSlicedNonbondedForceKernel


def _createKernelImpl(name, platform, context):
    if name == CalcSlicedNonbondedForceKernel.Name():
        return B200CalcSlicedNonbondedForceKernel(platform._properties.get("DeviceIndex", 0))
    raise OpenMMException("Tried to create kernel with illegal kernel name '%s'" % name)


def registerPlatforms():
    if "B200" not in Platform._platforms:
        Platform.registerPlatform(Platform("B200"))


def registerKernelFactories():
    registerPlatforms()
    Platform.getPlatformByName("B200").registerKernelFactory(CalcSlicedNonbondedForceKernel.Name(), _createKernelImpl)
