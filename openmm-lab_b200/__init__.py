"""B200-native SlicedNonbondedForce kernel: host-side mirror of the reference interface.

The compute path is the CUDA library in csrc/ behind the C-ABI of include/slicednb.h.
Importing this package does not load it; creating a kernel does, and fails loudly if the
library has not been built (there is no CPU fallback).
"""
from . import _abi
from ._abi import SnbError
from .context import Context, Platform, State, System
from .force import OpenMMException, SlicedNonbondedForce
from .kernel import B200CalcSlicedNonbondedForceKernel, CalcSlicedNonbondedForceKernel, load_library, partition
from .platform import registerKernelFactories, registerPlatforms
from . import systems
from .serialization import SlicedNonbondedForceProxy, XmlSerializer

registerKernelFactories()

__all__ = ["SlicedNonbondedForce", "System", "Context", "State", "Platform", "OpenMMException", "SnbError",
           "CalcSlicedNonbondedForceKernel", "B200CalcSlicedNonbondedForceKernel", "load_library",
           "registerKernelFactories", "registerPlatforms", "XmlSerializer", "SlicedNonbondedForceProxy"]
