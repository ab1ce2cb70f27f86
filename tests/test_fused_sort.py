"""The spatial sort as ONE cooperative kernel (sort.cu, k_sortFused) against the separate kernels it replaces.

The sorted order is a function of the positions only (cells along the Hilbert curve, atoms of a cell by index), so both
ways must produce the SAME order, hence the same tile list and -- forces being accumulated in fixed point -- bit-identical
direct-space forces.  It takes the launch for periodic systems of up to SORT_FUSED_MAX_CELLS = 4 096 cells (16 384 atoms) without
list reuse; above that the separate kernels are as fast (measured at DHFR size: 1.5 % faster).
SNB_FUSED_SORT=0 (read when a handle is created) selects the separate kernels; counter [0] tells which way it went."""
import os

import numpy as np
import pytest

import openmmlab_b200 as mm
from openmmlab_b200 import _abi, systems

pytestmark = pytest.mark.gpu


def _context(w, fused, props=None):
    old = os.environ.get("SNB_FUSED_SORT")
    os.environ["SNB_FUSED_SORT"] = "1" if fused else "0"
    try:
        ctx = mm.Context(w.system, "B200", dict(props or {}))
    finally:
        if old is None:
            del os.environ["SNB_FUSED_SORT"]
        else:
            os.environ["SNB_FUSED_SORT"] = old
    ctx.setPositions(w.positions)
    if w.box is not None:
        ctx.setPeriodicBoxVectors(*w.box)
    for k, v in w.parameters.items():
        ctx.setParameter(k, v)
    return ctx, ctx._kernels[0][1]


def _run(ctx, kernel, n, direct=True, reciprocal=True):
    ctx._forces = np.zeros((n, 3))
    ctx._energyParameterDerivatives = {}
    e = kernel.execute(ctx, True, True, direct, reciprocal)
    return e, ctx._forces.copy(), kernel.sliceEnergies.copy()


WORKLOADS = {
    "water_pme": lambda: systems.water_box(),
    "water_cutoff": lambda: systems.water_box(method=mm.SlicedNonbondedForce.CutoffPeriodic),
    "fep_ljpme": lambda: systems.fep_ligand(),
    "small_mixed": lambda: systems.small_mixed(n_side=7, solute_sites=16),
}


@pytest.mark.parametrize("name", sorted(WORKLOADS))
@pytest.mark.parametrize("props", [{}, {"DirectKernel": "packed"}, {"Precision": "double"}], ids=["general", "packed", "double"])
def test_fused_sort_matches_separate_kernels(name, props):
    w = WORKLOADS[name]()
    n = w.num_particles
    ctx_f, k_f = _context(w, True, props)
    ctx_s, k_s = _context(w, False, props)
    for rep in range(3):                # eager, captured, replayed
        e_f, f_f, s_f = _run(ctx_f, k_f, n, reciprocal=False)
        e_s, f_s, s_s = _run(ctx_s, k_s, n, reciprocal=False)
        assert k_f.getCounters()[0] == 1, "the fused sort kernel did not take the launch"
        assert k_s.getCounters()[0] == 0
        assert k_f.getCounters()[1] == k_s.getCounters()[1], "tile lists of different length"
        if props.get("Precision") == "double":
            assert np.abs(f_f-f_s).max() <= 1e-12*max(1.0, np.abs(f_s).max())
        else:
            assert np.array_equal(f_f, f_s), "rep %d: direct-space forces differ (max %g)" % (rep, np.abs(f_f-f_s).max())
        assert np.allclose(s_f, s_s, rtol=0, atol=1e-9*max(1.0, np.abs(s_s).max()))      # FP64 atomics: order-dependent last bits
    assert np.array_equal(k_f.dumpPairs(), k_s.dumpPairs())
    # the whole evaluation (reciprocal space spreads with float atomics: run-to-run noise in the last bits)
    e_f, f_f, s_f = _run(ctx_f, k_f, n)
    e_s, f_s, s_s = _run(ctx_s, k_s, n)
    rms = np.sqrt((f_s**2).sum(axis=1).mean())
    tol = 1e-10 if props.get("Precision") == "double" else 2e-6
    assert np.abs(f_f-f_s).max() <= tol*rms
    assert abs(e_f-e_s) <= tol*max(1.0, np.abs(s_s).sum())


def test_fused_sort_on_the_device_entry_point():
    import torch
    w = systems.fep_ligand()
    n = w.num_particles
    posq = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    posq[:, :3] = torch.from_numpy(w.positions.astype(np.float32)).cuda()
    box = np.ascontiguousarray(w.box, dtype=np.float64).reshape(9)
    out = {}
    for fused in (True, False):
        ctx, kernel = _context(w, fused, {"DeterministicForces": "true"})
        params = np.array([ctx.getParameter(name) for name in kernel.globalNames], dtype=np.float64)
        for rep in range(3):
            acc = torch.zeros((3, n), dtype=torch.int64, device="cuda")
            kernel.executeDevice(posq.data_ptr(), box, params, _abi.EXECUTE_ALL, acc.data_ptr())
            torch.cuda.synchronize()
            assert kernel.getCounters()[0] == (1 if fused else 0)
            got = acc.cpu().numpy(), kernel.readEnergies(_abi.EXECUTE_ALL)[0]
            if rep:
                assert np.array_equal(got[0], out[fused][0]), "replay differs from the eager evaluation"
            out[fused] = got
        # one launch for the sort against three (k_assignCells, k_scanScatter, k_boundsRank)
        launches = kernel.getCounters()[2]
        out[fused] += (launches,)
    assert np.array_equal(out[True][0], out[False][0]), "forces differ between the fused sort and the separate kernels"
    assert abs(out[True][1]-out[False][1]) <= 1e-12*max(1.0, abs(out[False][1]))
    assert out[False][2]-out[True][2] == 2


def test_list_reuse_and_large_systems_keep_the_separate_kernels():
    w = systems.water_box()
    ctx, kernel = _context(w, True, {"ListSkin": "0.05"})
    _run(ctx, kernel, w.num_particles)
    assert kernel.getCounters()[0] == 0
    w = systems.dhfr_like()             # 5 832 cells
    ctx, kernel = _context(w, True)
    _run(ctx, kernel, w.num_particles)
    assert kernel.getCounters()[0] == 0


@pytest.mark.parametrize("entry", ["host", "device"])
def test_begin_kernel_matches_copy_memset_convert(entry):
    """k_begin (parameters from the pinned mirror, accumulators zeroed, positions converted: one launch) against the copy,
    memset and conversion kernel it replaces (SNB_BEGIN_KERNEL=0), and k_finish's own stores of the energies / counters
    against the two device->host copies (SNB_HOST_RESULTS=0)."""
    import torch
    w = systems.small_mixed(n_side=8, solute_sites=20, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"])
    n = w.num_particles
    results = []
    for begin, host_results in (("1", "1"), ("0", "0")):
        os.environ["SNB_BEGIN_KERNEL"], os.environ["SNB_HOST_RESULTS"] = begin, host_results
        try:
            ctx, kernel = _context(w, True, {"DeterministicForces": "true"})
        finally:
            del os.environ["SNB_BEGIN_KERNEL"], os.environ["SNB_HOST_RESULTS"]
        for lam in (0.5, 0.25, 0.25):       # eager, new parameters (captured), replayed
            ctx.setParameter("lam", lam)
            if entry == "host":
                e, f, slices = _run(ctx, kernel, n)
            else:
                posq = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
                posq[:, :3] = torch.from_numpy(w.positions.astype(np.float32)).cuda()
                box = np.ascontiguousarray(w.box, dtype=np.float64).reshape(9)
                params = np.array([ctx.getParameter(name) for name in kernel.globalNames], dtype=np.float64)
                acc = torch.zeros((3, n), dtype=torch.int64, device="cuda")
                kernel.executeDevice(posq.data_ptr(), box, params, _abi.EXECUTE_ALL, acc.data_ptr())
                torch.cuda.synchronize()
                e, slices, _ = kernel.readEnergies(_abi.EXECUTE_ALL)
                f = acc.cpu().numpy().T/2.0**32
            results.append((begin, lam, e, f, slices.copy(), kernel.getCounters()[1]))
    half = len(results)//2
    for new, old in zip(results[:half], results[half:]):
        assert new[1] == old[1]
        assert np.array_equal(new[3], old[3]), "forces differ (lambda %g)" % new[1]
        assert abs(new[2]-old[2]) <= 1e-12*max(1.0, abs(old[2]))
        assert np.allclose(new[4], old[4], rtol=0, atol=1e-12*max(1.0, np.abs(old[4]).max()))
        assert new[5] == old[5] and new[5] > 0      # the list length came back through k_finish's own store


@pytest.mark.parametrize("name", ["dhfr", "water_pme"])
def test_compact_sort_matches_the_five_kernels(name):
    """k_assignCells -> k_scanScatter -> k_boundsRank (scan in shared memory, rank settled where the blocks are cut) against
    the five kernels with k_scanCells and k_rankInCell as launches of their own (SNB_COMPACT_SORT=0): same order, same list,
    bit-identical direct-space forces."""
    w = systems.dhfr_like() if name == "dhfr" else systems.water_box()
    n = w.num_particles
    out = {}
    for compact in ("1", "0"):
        os.environ["SNB_COMPACT_SORT"] = compact
        try:
            ctx, kernel = _context(w, False)
        finally:
            del os.environ["SNB_COMPACT_SORT"]
        for rep in range(3):
            e, f, slices = _run(ctx, kernel, n, reciprocal=False)
            assert kernel.getCounters()[0] == 0
        out[compact] = (f, kernel.getCounters()[1], kernel.getCounters()[2], kernel.dumpPairs())
    assert out["1"][1] == out["0"][1]
    assert np.array_equal(out["1"][0], out["0"][0])
    assert np.array_equal(out["1"][3], out["0"][3])
    assert out["0"][2]-out["1"][2] == 2


@pytest.mark.parametrize("name", ["water_pme", "fep_ljpme"])
def test_interpolation_folds_the_lambda_combination(name):
    """One or two subsets: k_pmeInterpolate combines the subsets' potential grids itself, on the points an atom reads
    (no k_pmeCombine pass over the whole grid); SNB_PME_FOLD=0 keeps the pass.  Same potentials, same forces."""
    w = WORKLOADS[name]()
    n = w.num_particles
    out = {}
    for fold in ("1", "0"):
        os.environ["SNB_PME_FOLD"] = fold
        try:
            ctx, kernel = _context(w, True, {"DeterministicForces": "true"})
            for rep in range(3):
                e, f, slices = _run(ctx, kernel, n, direct=False)
            out[fold] = (e, f, slices, kernel.getCounters()[2])
        finally:
            del os.environ["SNB_PME_FOLD"]
    terms = 2 if name == "fep_ljpme" else 1
    assert out["0"][3]-out["1"][3] == terms
    rms = np.sqrt((out["0"][1]**2).sum(axis=1).mean())
    assert np.abs(out["1"][1]-out["0"][1]).max() <= 1e-6*rms
    assert np.allclose(out["1"][2], out["0"][2], rtol=0, atol=1e-9*max(1.0, np.abs(out["0"][2]).max()))
