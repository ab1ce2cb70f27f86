"""Host logic of the CUDA core against the CPU oracle, without a GPU: what snb_create derives from a descriptor --
Ewald / PME / LJPME separation parameters and grids (OpenMM's calcPMEParameters rule), classic-Ewald reciprocal vectors
per axis (calcEwaldParameters) and the dispersion-correction coefficient of every slice
(SlicedNonbondedForceImpl::calcDispersionCorrections, openmmapi/src/SlicedNonbondedForceImpl.cpp:263-354) -- through the
host-only C-ABI entry point snb_host_parameters."""
import ctypes as C

import numpy as np
import pytest

import openmmlab_b200 as mm
from openmmlab_b200 import SlicedNonbondedForce as F, _abi, systems
from oracle import oracle_kernel

from util import CPU_PLATFORMS


def _host_parameters(w):
    lib = mm.load_library()
    packed = _abi.PackedDesc(w.force, np.asarray(w.system.getDefaultPeriodicBoxVectors(), dtype=np.float64))
    alpha, dalpha = C.c_double(), C.c_double()
    grid, dgrid, kmax = (C.c_int32*3)(), (C.c_int32*3)(), (C.c_int32*3)()
    coef = np.zeros(w.force.getNumSlices())
    fn = lib.snb_host_parameters
    fn.argtypes = [C.POINTER(_abi.SnbDesc), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int32),
                   C.POINTER(C.c_int32), C.c_void_p]
    assert fn(C.byref(packed.desc), C.byref(alpha), grid, C.byref(dalpha), dgrid, kmax, coef.ctypes.data) == _abi.OK
    return alpha.value, tuple(grid), dalpha.value, tuple(dgrid), tuple(kmax), coef


def _oracle(w, platform):
    ctx = mm.Context(w.system, platform)
    return ctx._kernels[0][1]


@pytest.mark.parametrize("platform", CPU_PLATFORMS)
@pytest.mark.parametrize("method", [F.CutoffPeriodic, F.Ewald, F.PME, F.LJPME])
def test_parameters_from_the_error_tolerance(platform, method):
    w = systems.small_mixed(method=method, num_groups=2, scaling=[("lam_a", 0.3, 0, 1, True, True)], derivatives=["lam_a"])
    w.force.setPMEParameters(0.0, 0, 0, 0)          # derive alpha and grid from the tolerance
    w.force.setLJPMEParameters(0.0, 0, 0, 0)
    w.force.setEwaldErrorTolerance(2e-4)
    w.force.setUseDispersionCorrection(True)
    alpha, grid, dalpha, dgrid, kmax, coef = _host_parameters(w)
    k = _oracle(w, platform)
    if method in (F.PME, F.LJPME):
        assert (alpha,)+grid == tuple(k.getPMEParameters())
    if method == F.LJPME:
        assert (dalpha,)+dgrid == tuple(k.getLJPMEParameters())
    if method == F.Ewald:
        assert (alpha,)+kmax == tuple(k.getEwaldParameters())
    np.testing.assert_allclose(coef, k.dispersionCoefficients(), rtol=1e-12, atol=0.0)
    if method != F.LJPME:
        assert np.abs(coef).max() > 0


@pytest.mark.parametrize("platform", CPU_PLATFORMS)
def test_dispersion_coefficients_of_the_bench_workloads(platform):
    # below 16 384 atoms the reference's integer arithmetic does not overflow: the 64-bit host code must agree
    for w in (systems.water_box(), systems.fep_ligand()):
        w.force.setUseDispersionCorrection(True)
        _, _, _, _, _, coef = _host_parameters(w)
        np.testing.assert_allclose(coef, _oracle(w, platform).dispersionCoefficients(), rtol=1e-12, atol=0.0)
