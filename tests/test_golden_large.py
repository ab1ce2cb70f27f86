"""Benchmark-size parity against the reference's own sources: the DHFR-, ApoA1- and STMV-size workloads of
BASELINE.json (configs[1], [2], [4]) checked on the GPU against fixtures generated in the build container from
oracle/_ref (tests/golden/make_golden_large.py) -- the reference's large-system tests are
tests/TestSlicedNonbondedForce.h:480-541 (testLargeSystem) and :543-598 (testHugeSystem).

Per workload: interacting-pair set bit-exact (count + SHA-256 of the sorted list); Precision=double within 1e-5 by the
reference's metric (AssertionUtilities.h:7-37) on total energy, every slice energy, every derivative and the forces
of a strided atom sample, for the direct, reciprocal and combined evaluation; the mixed mode -- both direct-space
kernels where both apply -- by util.assert_parity_mixed.  Nothing here reads /root/reference at run time.
"""
import hashlib
import os

import numpy as np
import pytest

from openmmlab_b200 import systems
from util import assert_parity, assert_parity_mixed, evaluate

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARTS = (("direct", dict(direct=True, reciprocal=False)), ("recip", dict(direct=False, reciprocal=True)),
         ("total", dict(direct=True, reciprocal=True)))


def _digest(a):
    h = hashlib.sha256()
    flat = np.ascontiguousarray(a).reshape(-1)
    for k in range(0, flat.size, 1 << 24):
        h.update(flat[k:k+(1 << 24)].tobytes())
    return h.hexdigest()


def load_large(w):
    z = np.load(os.path.join(GOLDEN, w.name+"_large.npz"))
    assert int(z["num_particles"]) == w.num_particles
    assert str(z["positions_sha256"]) == _digest(w.positions), "workload generator no longer reproduces the fixture's inputs"
    names = [str(s) for s in z["deriv_names"]]
    parts = {}
    for part, _ in PARTS:
        parts[part] = {"energy": float(z[part+"_energy"]), "slices": z[part+"_slices"], "forces": z[part+"_forces_sample"],
                       "derivs": dict(zip(names, z[part+"_derivs"])), "rms_force": float(z[part+"_rms_force"])}
    return z, parts, int(z["sample_stride"])


def sampled(got, stride):
    out = dict(got)
    out["forces"] = got["forces"][::stride]
    return out


def check_pairs(z, kernel, full):
    assert int(z["pair_count"]) == kernel.countPairs()
    if full:
        pairs = kernel.dumpPairs()
        assert str(z["pair_sha256"]) == _digest(pairs.astype(np.int32, copy=False)), "interacting pair set differs from the reference's"


@pytest.mark.parametrize("name", ["dhfr", "apoa1", "stmv"])
def test_large_workload_matches_reference(name):
    w = systems.WORKLOADS[name]()
    z, parts, stride = load_large(w)
    two_kernels = w.force.getNumSubsets() <= 2
    # ---- Precision=double: the reference's metric at 1e-5 on everything
    for part, flags in PARTS:
        got = evaluate("B200", w, properties={"Precision": "double"}, **flags)
        if part == "direct":
            check_pairs(z, got["kernel"], full=True)
        assert_parity(parts[part], sampled(got, stride), 1e-5, "%s %s [double]" % (w.name, part))
        del got
    # ---- mixed: both direct-space kernels
    for kernel in (("general", "packed") if two_kernels else ("auto",)):
        for part, flags in PARTS:
            if part == "recip" and kernel == "packed":
                continue
            got = evaluate("B200", w, properties={"DirectKernel": kernel}, **flags)
            if part == "direct":
                check_pairs(z, got["kernel"], full=True)
            what = "%s %s [mixed, DirectKernel=%s]" % (w.name, part, kernel)
            if part == "total":
                assert_parity_mixed(parts[part], sampled(got, stride), parts["direct"], parts["recip"], what=what, energy_tol=1e-5)
            else:
                assert_parity_mixed(parts[part], sampled(got, stride), what=what, energy_tol=1e-5)
            del got
