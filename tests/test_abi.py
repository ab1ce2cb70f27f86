"""The drop-in boundary on CPU: the C-ABI library loads without a GPU and exports every entry point
include/slicednb.h declares (no compute calls here), host-only entry points behave, and the C++ OpenMM
plugin adapter (plugin/) compiles against the reference's own headers where they are available."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "slicednb.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(snb_[a-z0-9_]+)\s*\(", text)
    return sorted(set(n for n in names if n != "snb_slice_index"))      # static inline helper, not exported


def test_library_exports_every_declared_symbol():
    import openmmlab_b200 as mm
    lib = mm.load_library()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "%s is declared in include/slicednb.h but not exported" % n


def test_oracle_exports_the_same_kernel_interface():
    from oracle import oracle_kernel
    oracle_kernel.build()
    lib = C.CDLL(oracle_kernel.PORT_LIB)
    for n in ("create", "destroy", "execute", "update_parameters", "get_pme_parameters", "get_ljpme_parameters", "last_error"):
        assert hasattr(lib, "snb_oracle_"+n)


def test_host_only_entry_points():
    import openmmlab_b200 as mm
    lib = mm.load_library()
    lib.snb_version.restype = C.c_char_p
    assert b"sm_100a" in lib.snb_version()
    lib.snb_last_error.restype = C.c_char_p
    assert isinstance(lib.snb_last_error(), bytes)
    # a descriptor of the wrong size is refused before any CUDA call
    desc = mm._abi.SnbDesc()
    desc.struct_size = 8
    handle = C.c_void_p()
    lib.snb_create.argtypes = [C.POINTER(mm._abi.SnbDesc), C.POINTER(C.c_void_p)]
    assert lib.snb_create(C.byref(desc), C.byref(handle)) == mm._abi.ERR_INVALID
    assert b"struct_size" in lib.snb_last_error()


def test_descriptor_layout_matches_header():
    # the ctypes mirror and the C struct must agree on the size (the library checks struct_size at run time)
    import openmmlab_b200 as mm
    src = '#include "slicednb.h"\n#include <stdio.h>\nint main(void) { printf("%zu", sizeof(snb_desc)); return 0; }\n'
    exe = os.path.join(ROOT, "tests", ".sizeof_desc")
    out = subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe], input=src, text=True, capture_output=True)
    assert out.returncode == 0, out.stderr
    try:
        size = int(subprocess.run([exe], capture_output=True, text=True).stdout)
    finally:
        os.remove(exe)
    assert size == C.sizeof(mm._abi.SnbDesc)


@pytest.mark.skipif(not os.path.isdir("/root/reference/openmmapi/include"), reason="the reference's headers are not on this machine")
def test_openmm_plugin_adapter_compiles_against_reference_headers():
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "plugin"), "-B", "check"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout+out.stderr
    obj = os.path.join(ROOT, "plugin", "build", "B200OpenMMLabKernels.check.o")
    syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
    for needed in ("registerKernelFactories", "registerPlatforms", "B200CalcSlicedNonbondedForceKernel::execute",
                   "B200CalcSlicedNonbondedForceKernel::initialize", "B200CalcSlicedNonbondedForceKernel::copyParametersToContext",
                   "B200OpenMMLabKernelFactory::createKernelImpl"):
        assert needed in syms
    # the zero-copy binding to OpenMM's CUDA platform (stand-in CudaContext headers): snb_execute_device_ordered is what it drives
    obj = os.path.join(ROOT, "plugin", "build", "B200CudaOpenMMLabKernels.check.o")
    syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
    for needed in ("registerOpenMMLabB200CudaKernelFactories", "B200CudaCalcSlicedNonbondedForceKernel::execute",
                   "B200CudaCalcSlicedNonbondedForceKernel::initialize", "B200CudaOpenMMLabKernelFactory::createKernelImpl",
                   "snb_execute_device_ordered", "snb_set_list_skin"):
        assert needed in syms


@pytest.mark.skipif(not os.path.isfile("/root/reference/openmmapi/src/SlicedNonbondedForce.cpp"), reason="the reference's sources are not on this machine")
@pytest.mark.parametrize("backend", ["ref", "port"])
def test_openmm_plugin_adapter_links_and_runs(backend):
    """plugin/Makefile run-check: the adapter, the reference's OWN openmmapi/src/SlicedNonbondedForce.cpp (compiled unchanged) and
    the stand-in OpenMM linked into one program that registers the factory through the plugin's registerKernelFactories(),
    creates the kernel by name and drives initialize / execute / copyParametersToContext / getPMEParameters against closed-form
    answers -- with the CPU oracle behind the C-ABI here (the CUDA library takes its place in `make run-check-cuda`)."""
    libdir, lib = ("../oracle/_ref", "libsnb_oracle_ref.so") if backend == "ref" else ("../oracle", "libsnb_oracle_port.so")
    if not os.path.exists(os.path.join(ROOT, "plugin", libdir, lib)):
        pytest.skip("%s is not built here" % lib)
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "plugin"), "-B", "run-check", "ORACLE_DIR="+libdir, "ORACLE_LIB="+lib],
                         capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().splitlines()[-2].strip() == "OK", out.stdout[-2000:]+out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isfile("/root/reference/tests/TestSlicedNonbondedForce.h"), reason="the reference's sources are not on this machine")
@pytest.mark.parametrize("backend", ["ref", "port"])
def test_reference_test_program_runs_through_the_adapter(backend):
    """plugin/Makefile run-reference-tests: the reference's OWN test program for SlicedNonbondedForce
    (tests/TestSlicedNonbondedForce.h with its main(): testCoulomb ... testDirectAndReciprocal, testOpenMMLab and
    testScalingParameterSeparation over all six methods), its own SlicedNonbondedForce.cpp and SlicedNonbondedForceImpl.cpp --
    all compiled UNCHANGED from /root/reference -- on the stand-in OpenMM of plugin/mock_openmm, through the plugin adapter,
    with the CPU oracle behind the C-ABI: every known-answer test of the reference pins the oracle (and the adapter) directly.
    (A plain NonbondedForce in those tests is evaluated as a one-subset SlicedNonbondedForce: OpenMM's own kernels are not here.)"""
    libdir, lib = ("../oracle/_ref", "libsnb_oracle_ref.so") if backend == "ref" else ("../oracle", "libsnb_oracle_port.so")
    if not os.path.exists(os.path.join(ROOT, "plugin", libdir, lib)):
        pytest.skip("%s is not built here" % lib)
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "plugin"), "-B", "run-reference-tests", "ORACLE_DIR="+libdir, "ORACLE_LIB="+lib],
                         capture_output=True, text=True)
    assert out.returncode == 0 and "\nDone\n" in out.stdout and "exception" not in out.stdout, out.stdout[-2000:]+out.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["mixed", "double"])
@pytest.mark.xfail(strict=False, reason="evidence only: linked in the build container, never yet run on a GPU when this test was written")
def test_reference_test_program_on_the_cuda_library(precision):
    """The same program linked against the CUDA library (make -C plugin build/run_reference_tests_cuda, done by build() where the
    reference's sources are; the binary travels to the GPU box).  Not a gate: it reports XPASS / XFAIL."""
    exe = os.path.join(ROOT, "plugin", "build", "run_reference_tests_cuda")
    if not os.path.exists(exe):
        pytest.skip("plugin/build/run_reference_tests_cuda was not built")
    env = dict(os.environ, SNB_PRECISION=precision)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "Done" in out.stdout, out.stdout[-2000:]+out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isfile("/root/reference/platforms/reference/src/ReferenceOpenMMLabKernels.cpp"), reason="the reference's sources are not on this machine")
def test_reference_own_kernel_passes_its_own_tests_on_the_stand_in_openmm():
    """plugin/Makefile run-reference-tests-own: the reference's own Reference-platform kernel (kernel glue, factory, pair loops, PME,
    API class and Impl: every source unchanged) under its own test program -- nothing of this repository on the path but the
    stand-in OpenMM, which this run validates (NonbondedForce as a one-subset force, PME / Ewald parameter rules, neighbour list)."""
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "plugin"), "-B", "run-reference-tests-own"], capture_output=True, text=True)
    assert out.returncode == 0 and "\nDone\n" in out.stdout and "exception" not in out.stdout, out.stdout[-2000:]+out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isfile("/root/reference/platforms/reference/src/ReferenceOpenMMLabKernels.cpp"), reason="the reference's sources are not on this machine")
@pytest.mark.parametrize("backend", ["ref", "port"])
def test_oracle_kernel_glue_matches_the_reference_own_glue(backend):
    """plugin/Makefile run-diff-own-vs-oracle: oracle/harness.cpp RESTATES the kernel glue of ReferenceOpenMMLabKernels.cpp (it
    cannot compile without OpenMM); here the real glue, compiled unchanged on the stand-in OpenMM, and the oracle (through the
    plugin adapter) evaluate the same random systems -- three subsets, exceptions from bonds, scaling parameters with derivatives,
    particle and exception offsets, switching function, dispersion correction, separate reciprocal-space force group, new
    parameter values, updateParametersInContext -- with all six methods, also in a triclinic box: 61 455 numbers equal to 1e-9 of their scale (measured: 4e-14)."""
    libdir, lib = ("../oracle/_ref", "libsnb_oracle_ref.so") if backend == "ref" else ("../oracle", "libsnb_oracle_port.so")
    if not os.path.exists(os.path.join(ROOT, "plugin", libdir, lib)):
        pytest.skip("%s is not built here" % lib)
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "plugin"), "-B", "run-diff-own-vs-oracle", "ORACLE_DIR="+libdir, "ORACLE_LIB="+lib],
                         capture_output=True, text=True)
    assert out.returncode == 0 and "\nOK\n" in out.stdout and "61455 comparisons" in out.stdout, out.stdout[-2000:]+out.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("double", "1e-7"), ("mixed", "1e-4")])
@pytest.mark.xfail(strict=False, reason="evidence only: linked in the build container, never yet run on a GPU when this test was written")
def test_reference_own_kernel_against_the_cuda_library(precision, tol):
    """plugin/build/diff_own_vs_cuda (built by build() where the reference's sources are): the reference's own Reference-platform
    kernel and the CUDA library behind the plugin adapter on the same random systems, six methods.  Not a gate."""
    exe = os.path.join(ROOT, "plugin", "build", "diff_own_vs_cuda")
    if not os.path.exists(exe):
        pytest.skip("plugin/build/diff_own_vs_cuda was not built")
    env = dict(os.environ, SNB_PRECISION=precision, DIFF_TOL=tol)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    print(out.stdout[-500:])
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout[-2000:]+out.stderr[-2000:]
