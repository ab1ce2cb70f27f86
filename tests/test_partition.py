"""Host-side logic of the multi-GPU path, on CPU: the partition rule exported by the C-ABI
(snb_partition, the one the tile-list build and the bonded kernel use) and the exact
fixed-point force reduction, exercised with world_size = 2 over gloo.

What the GPUs do per rank -- evaluate the pairs of a contiguous range of i units, convert the
forces to 2^32 fixed point, sum over ranks -- is emulated in numpy for the reaction-field method of
the 648-atom water box, and the reduced result is compared with the CPU oracle and with the
single-rank result (bit-identical: integer sums do not depend on the partition).
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ONE_4PI_EPS0 = 138.93545764438198


@pytest.mark.parametrize("n", [0, 1, 7, 2948, 133332])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("share", [0.0, 0.25, 1.0, 2.5])
def test_partition_covers_range(n, world, share):
    import openmmlab_b200 as mm
    edges = [mm.partition(n, world, r, share) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for r in range(world):
        lo, hi = edges[r]
        assert 0 <= lo <= hi <= n
        if r:
            assert lo == edges[r-1][1]
    if world > 1 and n > 100:
        sizes = [hi-lo for lo, hi in edges]
        others = sizes[1:]
        assert max(others)-min(others) <= 1
        if share == 0.0:
            assert sizes[0] == 0
        else:
            assert abs(sizes[0]-share*np.mean(others)) <= 1.0+share


def test_partition_rejects_bad_arguments():
    import openmmlab_b200 as mm
    with pytest.raises(mm.SnbError):
        mm.partition(10, 2, 2, 1.0)
    with pytest.raises(mm.SnbError):
        mm.partition(-1, 2, 0, 1.0)


def _partial_fixed_forces(w, pairs, lo, hi):
    """Reaction-field Coulomb + LJ forces (ReferenceSlicedLJCoulombIxn.cpp:553-611) of the pairs whose
    i unit (8 consecutive atoms) lies in [lo, hi), as [3][N] int64 in 2^32 fixed point."""
    f = w.force
    n = w.num_particles
    prm = np.asarray(f._particles, dtype=np.float64)
    q, sig, eps = prm[:, 0], 0.5*prm[:, 1], 2.0*np.sqrt(prm[:, 2])
    subset = np.zeros(n, dtype=np.int64)
    for k, v in f._subsets.items():
        subset[k] = v
    lam = np.ones((f.getNumSlices(), 2))
    for k in range(f.getNumScalingParameters()):
        name, s1, s2, coul, lj = f.getScalingParameter(k)
        sl = max(s1, s2)*(max(s1, s2)+1)//2+min(s1, s2)
        if coul:
            lam[sl, 0] = w.parameters[name]
        if lj:
            lam[sl, 1] = w.parameters[name]
    unit = np.minimum(pairs[:, 0], pairs[:, 1])//8
    mine = pairs[(unit >= lo) & (unit < hi)]
    i, j = mine[:, 0], mine[:, 1]
    L = np.diag(np.asarray(w.box))
    d = w.positions[j]-w.positions[i]
    d -= L*np.floor(d/L+0.5)
    r2 = (d*d).sum(axis=1)
    r = np.sqrt(r2)
    rc, eps_rf = f.getCutoffDistance(), f.getReactionFieldDielectric()
    krf = rc**-3*(eps_rf-1.0)/(2.0*eps_rf+1.0)
    s6 = ((sig[i]+sig[j])/r)**6
    e = eps[i]*eps[j]
    a, b = np.maximum(subset[i], subset[j]), np.minimum(subset[i], subset[j])
    sl = a*(a+1)//2+b
    dEdR = lam[sl, 1]*e*(12.0*s6-6.0)*s6/r2+lam[sl, 0]*ONE_4PI_EPS0*q[i]*q[j]*(1.0/r-2.0*krf*r2)/r2
    fij = np.rint(dEdR[:, None]*d*2.0**32).astype(np.int64)     # force on j; -fij on i
    out = np.zeros((3, n), dtype=np.int64)
    for c in range(3):
        np.add.at(out[c], j, fij[:, c])
        np.add.at(out[c], i, -fij[:, c])
    return out


def _worker(rank, world, port, share, queue):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import openmmlab_b200 as mm
    from openmmlab_b200 import SlicedNonbondedForce as F
    from openmmlab_b200 import systems
    from util import evaluate
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        w = systems.water_box(method=F.CutoffPeriodic, dispersion_correction=False)
        ref = evaluate("ReferencePort", w, pairs=True, reciprocal=False)
        units = (w.num_particles+7)//8
        lo, hi = mm.partition(units, world, rank, share)
        ranges = [None]*world
        dist.all_gather_object(ranges, (lo, hi))
        assert ranges[0][0] == 0 and ranges[-1][1] == units and all(ranges[r][0] == ranges[r-1][1] for r in range(1, world))
        part = torch.from_numpy(_partial_fixed_forces(w, ref["pairs"], lo, hi))
        dist.reduce(part, dst=0, op=dist.ReduceOp.SUM)          # what ncclReduce(int64, sum) does on the GPUs
        if rank == 0:
            whole = _partial_fixed_forces(w, ref["pairs"], 0, units)
            reduced = part.numpy()
            forces = reduced.T/2.0**32
            queue.put((bool(np.array_equal(reduced, whole)), float(np.abs(forces-ref["forces"]).max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("share", [1.0, 0.25, 0.0])
def test_partitioned_fixed_point_reduce_gloo(share):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, share, queue)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    identical, err = queue.get()
    assert identical, "the reduced fixed-point forces depend on the partition"
    assert err < 1e-4, "reduced forces differ from the oracle's by %g" % err
