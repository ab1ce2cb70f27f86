"""ctypes mirror of include/slicednb.h (the C-ABI drop-in boundary).

Only declarations live here: the descriptor struct, the enums and a helper that
flattens a :class:`SlicedNonbondedForce` (force.py) into an ``snb_desc``.  The
same descriptor is consumed by the product library (csrc/, CUDA) and by the CPU
oracle (oracle/, test infrastructure) so parity tests feed both byte-identical
inputs.
"""
import ctypes as C

import numpy as np

# NonbondedMethod -- openmmapi/include/OpenMMLabKernels.h:31-38
NoCutoff, CutoffNonPeriodic, CutoffPeriodic, Ewald, PME, LJPME = range(6)

INCLUDE_FORCES, INCLUDE_ENERGY, INCLUDE_DIRECT, INCLUDE_RECIPROCAL = 1, 2, 4, 8
FLAG_DETERMINISTIC, FLAG_DOUBLE_PRECISION, FLAG_DIRECT_PACKED, FLAG_DIRECT_GENERAL = 1, 2, 4, 8
EXECUTE_ALL = 15

OK, ERR_INVALID, ERR_BOX_TOO_SMALL, ERR_PARTICLE_COUNT, ERR_EXCEPTION_SET, \
    ERR_WRONG_METHOD, ERR_CUDA, ERR_UNSUPPORTED = range(8)

PHASES = ("sort", "list", "direct", "bonded", "spread", "fft", "conv", "interp", "reduce", "total")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class SnbDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("num_particles", C.c_int32),
        ("num_subsets", C.c_int32),
        ("method", C.c_int32),
        ("cutoff", C.c_double),
        ("use_switching_function", C.c_int32),
        ("use_dispersion_correction", C.c_int32),
        ("switching_distance", C.c_double),
        ("reaction_field_dielectric", C.c_double),
        ("exceptions_use_pbc", C.c_int32),
        ("flags", C.c_int32),
        ("ewald_error_tolerance", C.c_double),
        ("default_box", C.c_double*9),
        ("ewald_alpha", C.c_double),
        ("pme_grid", C.c_int32*3),
        ("device_index", C.c_int32),
        ("dispersion_alpha", C.c_double),
        ("dispersion_grid", C.c_int32*3),
        ("reserved0", C.c_int32),
        ("charge", _dp),
        ("sigma", _dp),
        ("epsilon", _dp),
        ("subset", _ip),
        ("num_exceptions", C.c_int32),
        ("num_global_parameters", C.c_int32),
        ("exception_particles", _ip),
        ("exception_params", _dp),
        ("global_parameter_defaults", _dp),
        ("num_particle_offsets", C.c_int32),
        ("num_exception_offsets", C.c_int32),
        ("particle_offset_index", _ip),
        ("particle_offset_scale", _dp),
        ("exception_offset_index", _ip),
        ("exception_offset_scale", _dp),
        ("num_scaling_parameters", C.c_int32),
        ("num_derivatives", C.c_int32),
        ("scaling_parameters", _ip),
        ("derivative_parameters", _ip),
    ]


class SnbError(RuntimeError):
    """Raised where the reference throws OpenMMException (status code kept in .code)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


def _f64(a, shape):
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(shape))
    return arr


def _i32(a, shape):
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(shape))
    return arr


class PackedDesc:
    """An ``snb_desc`` plus the numpy arrays that keep its pointers alive."""

    def __init__(self, force, default_box=None, device_index=0, flags=0):
        n = force.getNumParticles()
        d = SnbDesc()
        keep = []

        def dptr(a, shape):
            arr = _f64(a, shape)
            keep.append(arr)
            return arr.ctypes.data_as(_dp)

        def iptr(a, shape):
            arr = _i32(a, shape)
            keep.append(arr)
            return arr.ctypes.data_as(_ip)

        d.struct_size = C.sizeof(SnbDesc)
        d.num_particles = n
        d.num_subsets = force.getNumSubsets()
        d.method = force.getNonbondedMethod()
        d.cutoff = force.getCutoffDistance()
        d.use_switching_function = int(force.getUseSwitchingFunction())
        d.use_dispersion_correction = int(force.getUseDispersionCorrection())
        d.switching_distance = force.getSwitchingDistance()
        d.reaction_field_dielectric = force.getReactionFieldDielectric()
        d.exceptions_use_pbc = int(force.getExceptionsUsePeriodicBoundaryConditions())
        d.flags = flags
        d.ewald_error_tolerance = force.getEwaldErrorTolerance()
        box = np.zeros(9) if default_box is None else np.asarray(default_box, dtype=np.float64).reshape(9)
        for k in range(9):
            d.default_box[k] = box[k]
        alpha, nx, ny, nz = force.getPMEParameters()
        d.ewald_alpha = alpha
        d.pme_grid[0], d.pme_grid[1], d.pme_grid[2] = nx, ny, nz
        alpha, nx, ny, nz = force.getLJPMEParameters()
        d.dispersion_alpha = alpha
        d.dispersion_grid[0], d.dispersion_grid[1], d.dispersion_grid[2] = nx, ny, nz
        d.device_index = device_index

        # bulk access to the container's own storage: the per-item getters are O(N) Python calls
        params = np.asarray(force._particles, dtype=np.float64).reshape(n, 3)
        d.charge = dptr(params[:, 0], (n,))
        d.sigma = dptr(params[:, 1], (n,))
        d.epsilon = dptr(params[:, 2], (n,))
        subsets = np.zeros(n, dtype=np.int32)
        if force._subsets:
            idx = np.fromiter(force._subsets.keys(), dtype=np.int64, count=len(force._subsets))
            subsets[idx] = np.fromiter(force._subsets.values(), dtype=np.int32, count=len(force._subsets))
        d.subset = iptr(subsets, (n,))

        x = force.getNumExceptions()
        exc = np.asarray(force._exceptions, dtype=np.float64).reshape(x, 5)
        d.num_exceptions = x
        d.exception_particles = iptr(exc[:, 0:2], (x, 2))
        d.exception_params = dptr(exc[:, 2:5], (x, 3))

        g = force.getNumGlobalParameters()
        d.num_global_parameters = g
        d.global_parameter_defaults = dptr([force.getGlobalParameterDefaultValue(i) for i in range(g)], (g,))
        names = [force.getGlobalParameterName(i) for i in range(g)]

        p = force.getNumParticleParameterOffsets()
        offs = [force.getParticleParameterOffset(i) for i in range(p)]
        d.num_particle_offsets = p
        d.particle_offset_index = iptr([[names.index(o[0]), o[1]] for o in offs], (p, 2))
        d.particle_offset_scale = dptr([[o[2], o[3], o[4]] for o in offs], (p, 3))
        e = force.getNumExceptionParameterOffsets()
        offs = [force.getExceptionParameterOffset(i) for i in range(e)]
        d.num_exception_offsets = e
        d.exception_offset_index = iptr([[names.index(o[0]), o[1]] for o in offs], (e, 2))
        d.exception_offset_scale = dptr([[o[2], o[3], o[4]] for o in offs], (e, 3))

        k = force.getNumScalingParameters()
        sps = [force.getScalingParameter(i) for i in range(k)]
        d.num_scaling_parameters = k
        d.scaling_parameters = iptr([[names.index(s[0]), s[1], s[2], int(s[3]), int(s[4])] for s in sps], (k, 5))
        nd = force.getNumScalingParameterDerivatives()
        dnames = [force.getScalingParameterDerivativeName(i) for i in range(nd)]
        d.num_derivatives = nd
        d.derivative_parameters = iptr([names.index(s) for s in dnames], (nd,))

        self.desc = d
        self.keep = keep
        self.global_names = names
        self.derivative_names = dnames
        self.num_slices = force.getNumSlices()


def as_f64_ptr(arr):
    return arr.ctypes.data_as(_dp)
