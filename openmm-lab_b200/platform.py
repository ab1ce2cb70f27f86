"""Kernel-factory registration for platform "B200".

Mirrors the reference's plugin entry points
(platforms/cuda/src/CudaOpenMMLabKernelFactory.cpp:19-57): ``registerKernelFactories``
registers a factory for the kernel name "CalcSlicedNonbondedForce"; the factory throws for
any other name (:56).  Platform properties follow the reference's CUDA platform
(CudaOpenMMLabKernels.cpp:483,496; platforms/cuda/tests/CudaOpenMMLabTests.h:26-31):
``Precision`` (mixed | double), ``DeterministicForces``, ``DeviceIndex``; plus ``DirectKernel``
(auto | packed | general: include/slicednb.h SNB_FLAG_DIRECT_*) and ``ListSkin`` (nm; > 0 keeps the
tile list between evaluations until an atom has moved half of it, as OpenMM's CUDA platform pads
and reuses its neighbour list; include/slicednb.h snb_set_list_skin).
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
        try:
            list_skin = float(props.get("ListSkin", 0.0))
        except ValueError:
            list_skin = -1.0
        if not 0.0 <= list_skin <= 10.0:
            raise OpenMMException("Illegal value for ListSkin: %s (nm, 0 = rebuild the tile list every evaluation)" % props.get("ListSkin"))
        return B200CalcSlicedNonbondedForceKernel(int(props.get("DeviceIndex", 0)), precision, deterministic, direct_kernel, list_skin)
    raise OpenMMException("Tried to create kernel with illegal kernel name '%s'" % name)


def registerPlatforms():
    if "B200" not in Platform._platforms:
        Platform.registerPlatform(Platform("B200"))


def registerKernelFactories():
    registerPlatforms()
    Platform.getPlatformByName("B200").registerKernelFactory(CalcSlicedNonbondedForceKernel.Name(), _createKernelImpl)
