"""Host-side mirror of the reference's kernel interface for this path.

``CalcSlicedNonbondedForceKernel`` has the five methods of the reference's abstract
class (openmmapi/include/OpenMMLabKernels.h:29-87) -- ``initialize``, ``execute``,
``copyParametersToContext``, ``getPMEParameters``, ``getLJPMEParameters`` -- each a
thin forward to one C-ABI entry point of include/slicednb.h.  The B200 kernel class
below binds them to the CUDA library built from csrc/; it has NO CPU fallback and
raises at import of the library if the extension is missing.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import PackedDesc, SnbDesc, SnbError

_dp = C.POINTER(C.c_double)


class CalcSlicedNonbondedForceKernel:
    """Binds the kernel interface to a C library exporting ``<prefix>create`` etc."""

    NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME, LJPME = range(6)

    @staticmethod
    def Name():
        return "CalcSlicedNonbondedForce"

    def __init__(self, lib, prefix, device_index=0, flags=0):
        self._lib = lib
        self._prefix = prefix
        self._handle = C.c_void_p()
        self._device_index = device_index
        self._flags = flags
        self._packed = None
        self._execute = None

    def _fn(self, name):
        return getattr(self._lib, self._prefix+name)

    def _check(self, status):
        if status != _abi.OK:
            err = self._fn("last_error")
            err.restype = C.c_char_p
            raise SnbError(status, err().decode())

    # initialize(system, force) -- OpenMMLabKernels.h:50
    def initialize(self, system, force):
        self.destroy()
        box = system.getDefaultPeriodicBoxVectors()
        self._packed = PackedDesc(force, np.asarray(box, dtype=np.float64), self._device_index, self._flags)
        create = self._fn("create")
        create.argtypes = [C.POINTER(SnbDesc), C.POINTER(C.c_void_p)]
        self._check(create(C.byref(self._packed.desc), C.byref(self._handle)))
        self.numParticles = force.getNumParticles()
        self.numSlices = force.getNumSlices()
        self.globalNames = self._packed.global_names
        self.derivativeNames = self._packed.derivative_names
        self.sliceEnergies = np.zeros((self.numSlices, 2))
        self._execute = None

    # execute(context, includeForces, includeEnergy, includeDirect, includeReciprocal) -- OpenMMLabKernels.h:61
    def execute(self, context, includeForces, includeEnergy, includeDirect, includeReciprocal):
        flags = (_abi.INCLUDE_FORCES if includeForces else 0) | (_abi.INCLUDE_ENERGY if includeEnergy else 0) | \
                (_abi.INCLUDE_DIRECT if includeDirect else 0) | (_abi.INCLUDE_RECIPROCAL if includeReciprocal else 0)
        pos = context._positions
        if pos.dtype != np.float64 or not pos.flags.c_contiguous:
            pos = np.ascontiguousarray(pos, dtype=np.float64)
        box = context._box
        if box.dtype != np.float64 or not box.flags.c_contiguous:
            box = np.ascontiguousarray(box, dtype=np.float64)
        fn = self._execute
        if fn is None:
            # (scratch arrays and the ctypes prototype are made once per initialize, not once per step)
            fn = self._execute = self._fn("execute")
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            self._exec_params = np.zeros(max(1, len(self.globalNames)))
            self._exec_derivs = np.zeros(max(1, len(self.derivativeNames)))
            self._exec_energy = np.zeros(1)
        params, derivs = self._exec_params, self._exec_derivs
        for k, name in enumerate(self.globalNames):
            params[k] = context.getParameter(name)
        derivs.fill(0.0)
        forces = context._forces
        status = fn(self._handle, pos.ctypes.data, box.ctypes.data, params.ctypes.data, flags,
                    forces.ctypes.data if includeForces else None, self._exec_energy.ctypes.data,
                    self.sliceEnergies.ctypes.data, derivs.ctypes.data)
        if status != _abi.OK:
            self._check(status)
        for k, name in enumerate(self.derivativeNames):
            context._energyParameterDerivatives[name] = context._energyParameterDerivatives.get(name, 0.0)+derivs[k]
        return float(self._exec_energy[0])

    # copyParametersToContext(context, force) -- OpenMMLabKernels.h:68
    def copyParametersToContext(self, context, force):
        packed = PackedDesc(force, np.asarray(context._system.getDefaultPeriodicBoxVectors(), dtype=np.float64), self._device_index, self._flags)
        fn = self._fn("update_parameters")
        fn.argtypes = [C.c_void_p, C.POINTER(SnbDesc)]
        self._check(fn(self._handle, C.byref(packed.desc)))
        self._packed = packed

    def _pme(self, name):
        alpha, nx, ny, nz = C.c_double(), C.c_int(), C.c_int(), C.c_int()
        fn = self._fn(name)
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self._check(fn(self._handle, C.byref(alpha), C.byref(nx), C.byref(ny), C.byref(nz)))
        return alpha.value, nx.value, ny.value, nz.value

    # getPMEParameters / getLJPMEParameters -- OpenMMLabKernels.h:77,86
    def getPMEParameters(self):
        return self._pme("get_pme_parameters")

    def getLJPMEParameters(self):
        return self._pme("get_ljpme_parameters")

    def dumpPairs(self):
        """Sorted (i<j) interacting pairs of the last execute (test introspection)."""
        ptr, count = C.POINTER(C.c_int32)(), C.c_int64()
        fn = self._fn("dump_pairs")
        fn.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int64)]
        self._check(fn(self._handle, C.byref(ptr), C.byref(count)))
        out = np.ctypeslib.as_array(ptr, shape=(max(count.value, 1), 2))[:count.value].copy()
        free = self._fn("free")
        free.argtypes = [C.c_void_p]
        free(ptr)
        return out

    def destroy(self):
        if self._handle:
            fn = self._fn("destroy")
            fn.argtypes = [C.c_void_p]
            fn.restype = None
            fn(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


_LIB = None
# SNB_B200_LIB: another build of the same CUDA core (kernel-tuning experiments, see build.py SNB_LIB_TAG)
LIB_PATH = os.environ.get("SNB_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libslicednb_b200.so")


def load_library():
    """Load the CUDA core.  Fails loudly: there is no CPU or PyTorch fallback for this path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python __graft_entry__.py build` "
                              "(nvcc, sm_100a); this path has no CPU fallback" % LIB_PATH)
        _LIB = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    return _LIB


class B200CalcSlicedNonbondedForceKernel(CalcSlicedNonbondedForceKernel):
    """The sm_100a implementation, registered on platform "B200" (platform.py)."""

    def __init__(self, device_index=0, precision="mixed", deterministic=False, direct_kernel="auto", list_skin=0.0):
        self.listSkin = float(list_skin)
        flags = (_abi.FLAG_DOUBLE_PRECISION if precision == "double" else 0) | (_abi.FLAG_DETERMINISTIC if deterministic else 0) | \
                {"auto": 0, "packed": _abi.FLAG_DIRECT_PACKED, "general": _abi.FLAG_DIRECT_GENERAL}[direct_kernel]
        super().__init__(load_library(), "snb_", device_index, flags)
        self.precision = precision
        self._device_fn = self._read_fn = None

    def initialize(self, system, force):
        super().initialize(system, force)
        self._read_fn = None            # (its scratch arrays are sized by the force's derivative list)
        if self.listSkin > 0:
            self.setListSkin(self.listSkin)

    def setListSkin(self, skin_nm):
        """Tile-list reuse (snb_set_list_skin): list built with cutoff + skin, kept until an atom has moved skin/2; 0 = off."""
        fn = self._lib.snb_set_list_skin
        fn.argtypes = [C.c_void_p, C.c_double]
        self._check(fn(self._handle, float(skin_nm)))
        self.listSkin = float(skin_nm)

    def executeDevice(self, posq_ptr, box, params, flags, forces_ptr):
        # (an MD loop calls this every step with the same arrays: the ctypes prototype is made once and the arrays are passed
        # by address -- a few microseconds saved per call, against an evaluation of ~150)
        fn = self._device_fn
        if fn is None:
            fn = self._device_fn = self._lib.snb_execute_device
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        if box.dtype != np.float64 or not box.flags.c_contiguous:
            box = np.ascontiguousarray(box, dtype=np.float64)
        if params.dtype != np.float64 or not params.flags.c_contiguous:
            params = np.ascontiguousarray(params, dtype=np.float64)
        status = fn(self._handle, posq_ptr, box.ctypes.data, params.ctypes.data, flags, forces_ptr)
        if status != _abi.OK:
            self._check(status)

    def executeDeviceOrdered(self, posq_ptr, box, params, flags, forces_ptr, padded_num_atoms, atom_index_ptr=None,
                             posq_correction_ptr=None, positions_double=False):
        """snb_execute_device_ordered: positions / forces in the caller's own (reordered, padded) atom order, as OpenMM's CUDA
        platform keeps them (include/slicednb.h, snb_device_layout)."""
        class Layout(C.Structure):
            _fields_ = [("struct_size", C.c_int32), ("padded_num_atoms", C.c_int32), ("positions_double", C.c_int32), ("reserved", C.c_int32),
                        ("posq", C.c_void_p), ("posq_correction", C.c_void_p), ("atom_index", C.c_void_p), ("forces", C.c_void_p)]
        layout = Layout(C.sizeof(Layout), int(padded_num_atoms), int(bool(positions_double)), 0, posq_ptr, posq_correction_ptr, atom_index_ptr, forces_ptr)
        fn = self._lib.snb_execute_device_ordered
        fn.argtypes = [C.c_void_p, C.c_void_p, _dp, _dp, C.c_int]
        box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
        params = np.ascontiguousarray(params, dtype=np.float64)
        self._check(fn(self._handle, C.byref(layout), box.ctypes.data_as(_dp), params.ctypes.data_as(_dp), flags))

    def readEnergies(self, flags):
        fn = self._read_fn
        if fn is None:
            fn = self._read_fn = self._lib.snb_read_energies
            fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            self._read_energy = np.zeros(1)
            self._read_derivs = np.zeros(max(1, len(self.derivativeNames)))
        self._read_derivs.fill(0.0)
        status = fn(self._handle, flags, self._read_energy.ctypes.data, self.sliceEnergies.ctypes.data, self._read_derivs.ctypes.data)
        if status != _abi.OK:
            self._check(status)
        return float(self._read_energy[0]), self.sliceEnergies.copy(), self._read_derivs[:len(self.derivativeNames)].copy()

    def setTiming(self, enabled):
        self._check(self._lib.snb_set_timing(self._handle, int(enabled)))

    def getPhaseTimes(self):
        ms = (C.c_float*len(_abi.PHASES))()
        self._check(self._lib.snb_get_phase_times(self._handle, ms))
        return dict(zip(_abi.PHASES, list(ms)))

    def countPairs(self):
        n = C.c_int64()
        fn = self._lib.snb_count_pairs
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        self._check(fn(self._handle, C.byref(n)))
        return n.value

    def getLastDeviceMs(self):
        ms = C.c_float()
        self._check(self._lib.snb_get_last_device_ms(self._handle, C.byref(ms)))
        return ms.value

    def setListCapacity(self, chunks):
        """Test hook (snb_set_list_capacity): too small a tile-list capacity exercises overflow -> grow -> repeat."""
        fn = self._lib.snb_set_list_capacity
        fn.argtypes = [C.c_void_p, C.c_int64]
        self._check(fn(self._handle, int(chunks)))

    def getCounters(self):
        out = (C.c_int64*8)()
        self._check(self._lib.snb_get_counters(self._handle, out))
        return list(out)

    # ---- multi-GPU: one process / one kernel per GPU (include/slicednb.h, "multi-GPU") ----
    def ncclUniqueId(self):
        """128-byte NCCL id made on rank 0; ship it to the other ranks (e.g. torch.distributed broadcast)."""
        buf = C.create_string_buffer(128)
        fn = self._lib.snb_nccl_unique_id
        fn.argtypes = [C.c_char_p]
        self._check(fn(buf))
        return buf.raw

    def commInit(self, unique_id, rank, world_size):
        fn = self._lib.snb_comm_init
        fn.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        self._check(fn(self._handle, unique_id, rank, world_size))
        self.rank, self.worldSize = rank, world_size

    def setCommOptions(self, broadcast_positions=True):
        fn = self._lib.snb_set_comm_options
        fn.argtypes = [C.c_void_p, C.c_int]
        self._check(fn(self._handle, int(bool(broadcast_positions))))

    def setRank0Share(self, share):
        fn = self._lib.snb_set_rank0_share
        fn.argtypes = [C.c_void_p, C.c_double]
        self._check(fn(self._handle, float(share)))


def partition(n, world_size, rank, rank0_share=1.0):
    """The library's partition rule (host only, no GPU needed): range [lo, hi) of n units for `rank`."""
    lib = load_library()
    lo, hi = C.c_int64(), C.c_int64()
    fn = lib.snb_partition
    fn.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    if fn(n, world_size, rank, float(rank0_share), C.byref(lo), C.byref(hi)) != _abi.OK:
        raise SnbError(_abi.ERR_INVALID, "bad partition arguments")
    return lo.value, hi.value
