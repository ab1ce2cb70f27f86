"""TEST INFRASTRUCTURE ONLY -- Python binding of the CPU oracle.

Registers two platforms that implement the reference's kernel interface
(openmmapi/include/OpenMMLabKernels.h:29-87) on the CPU, through the same descriptor as the
product's C-ABI:
  "Reference"      oracle/_ref/libsnb_oracle_ref.so -- the reference's OWN Reference-platform
                   sources compiled unchanged (oracle/Makefile `ref`), when that file exists;
                   otherwise the port below (``backend()`` says which).
  "ReferencePort"  oracle/libsnb_oracle_port.so -- the from-scratch restatement.
"""
import ctypes as C
import os
import subprocess
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(_here)
if _root not in sys.path:
    sys.path.insert(0, _root)

import openmmlab_b200  # noqa: E402  (host-side mirror: descriptor packing, Context, Platform)
from openmmlab_b200.context import Platform  # noqa: E402
from openmmlab_b200.kernel import CalcSlicedNonbondedForceKernel  # noqa: E402

PORT_LIB = os.path.join(_here, "libsnb_oracle_port.so")
REF_LIB = os.path.join(_here, "_ref", "libsnb_oracle_ref.so")
_libs = {}


def build(verbose=False):
    """Compile the oracle (port always; _ref when /root/reference is present)."""
    out = subprocess.run(["make", "-C", _here, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n"+out.stdout+out.stderr)
    if verbose:
        print(out.stdout)


def _load(path):
    if path not in _libs:
        if not os.path.exists(path):
            if path == PORT_LIB:
                build()
            else:
                raise ImportError(path+" is not built")
        _libs[path] = C.CDLL(path)
    return _libs[path]


def have_ref():
    return os.path.exists(REF_LIB)


class OracleKernel(CalcSlicedNonbondedForceKernel):
    def __init__(self, which="auto"):
        path = REF_LIB if (which == "ref" or (which == "auto" and have_ref())) else PORT_LIB
        super().__init__(_load(path), "snb_oracle_")
        self._lib.snb_oracle_backend.restype = C.c_char_p
        self._lib.snb_oracle_cutoff_margin.restype = C.c_double

    def backend(self):
        return self._lib.snb_oracle_backend().decode()

    def brutePairs(self, positions, box):
        import numpy as np
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
        ptr, count = C.POINTER(C.c_int32)(), C.c_int64()
        dp = C.POINTER(C.c_double)
        fn = self._lib.snb_oracle_brute_pairs
        fn.argtypes = [C.c_void_p, dp, dp, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int64)]
        self._check(fn(self._handle, pos.ctypes.data_as(dp), box.ctypes.data_as(dp), C.byref(ptr), C.byref(count)))
        out = np.ctypeslib.as_array(ptr, shape=(max(count.value, 1), 2))[:count.value].copy()
        self._lib.snb_oracle_free.argtypes = [C.c_void_p]
        self._lib.snb_oracle_free(ptr)
        return out

    def cutoffMargin(self, positions, box):
        import numpy as np
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
        dp = C.POINTER(C.c_double)
        fn = self._lib.snb_oracle_cutoff_margin
        fn.argtypes = [C.c_void_p, dp, dp]
        return fn(self._handle, pos.ctypes.data_as(dp), box.ctypes.data_as(dp))

    def dispersionCoefficients(self):
        import numpy as np
        out = np.zeros(self.numSlices)
        fn = self._lib.snb_oracle_dispersion_coefficients
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self._check(fn(self._handle, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def getEwaldParameters(self):
        alpha, kx, ky, kz = C.c_double(), C.c_int(), C.c_int(), C.c_int()
        fn = self._lib.snb_oracle_get_ewald_parameters
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self._check(fn(self._handle, C.byref(alpha), C.byref(kx), C.byref(ky), C.byref(kz)))
        return alpha.value, kx.value, ky.value, kz.value


def register():
    for name, which in (("Reference", "auto"), ("ReferencePort", "port")):
        if name not in Platform._platforms:
            Platform.registerPlatform(Platform(name))
        Platform.getPlatformByName(name).registerKernelFactory(
            CalcSlicedNonbondedForceKernel.Name(), lambda n, p, c, w=which: OracleKernel(w))
    if have_ref():
        if "ReferenceRef" not in Platform._platforms:
            Platform.registerPlatform(Platform("ReferenceRef"))
        Platform.getPlatformByName("ReferenceRef").registerKernelFactory(
            CalcSlicedNonbondedForceKernel.Name(), lambda n, p, c: OracleKernel("ref"))


register()
