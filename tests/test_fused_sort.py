"""The spatial sort as ONE cooperative kernel (sort.cu, k_sortFused) against the separate kernels it replaces.

The sorted order is a function of the positions only (cells along the Hilbert curve, atoms of a cell by index), so both
ways must produce the SAME order, hence the same tile list and -- forces being accumulated in fixed point -- bit-identical
direct-space forces; the fused kernel also takes over the position conversion of a device-resident caller.
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
    "dhfr": lambda: systems.dhfr_like(),
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


def test_fused_sort_converts_device_positions():
    import torch
    w = systems.dhfr_like()
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
        # one launch for the sort and none for the conversion, against five and one
        launches = kernel.getCounters()[2]
        out[fused] += (launches,)
    assert np.array_equal(out[True][0], out[False][0]), "forces differ between the fused sort and the separate kernels"
    assert abs(out[True][1]-out[False][1]) <= 1e-12*max(1.0, abs(out[False][1]))
    assert out[False][2]-out[True][2] == 5


def test_list_reuse_and_nonperiodic_keep_the_separate_kernels():
    w = systems.water_box()
    ctx, kernel = _context(w, True, {"ListSkin": "0.05"})
    _run(ctx, kernel, w.num_particles)
    assert kernel.getCounters()[0] == 0
