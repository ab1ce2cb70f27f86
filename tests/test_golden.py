"""Golden fixtures (tests/golden/*.npz): outputs of the reference's OWN Reference-platform sources
(oracle/_ref, compiled unchanged from /root/reference in the build container by
tests/golden/make_golden.py).  They pin

  * the oracle port (CPU, unmarked): 1e-9 by the reference's metric -- the port and the reference
    sources are both double precision and differ only in summation order and libm calls;
  * the CUDA path (platform "B200", gpu): Precision=double at 1e-5 on every quantity, and the default
    mixed mode by util.assert_parity_mixed; interacting-pair sets bit-exact (count + SHA-256 of the
    sorted list).

Nothing here reads /root/reference at run time.
"""
import hashlib
import os

import numpy as np
import pytest

from golden_cases import CASES
from util import assert_parity, assert_parity_mixed, evaluate

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PARTS = (("direct", dict(direct=True, reciprocal=False)), ("recip", dict(direct=False, reciprocal=True)),
         ("total", dict(direct=True, reciprocal=True)))


def _load(name):
    z = np.load(os.path.join(GOLDEN, name+".npz"))
    names = [str(s) for s in z["deriv_names"]]
    parts = {}
    for part, _ in PARTS:
        parts[part] = {"energy": float(z[part+"_energy"]), "slices": z[part+"_slices"], "forces": z[part+"_forces"],
                       "derivs": dict(zip(names, z[part+"_derivs"]))}
    return z, parts


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _check_inputs(z, w):
    assert int(z["num_particles"]) == w.num_particles
    assert str(z["positions_sha256"]) == _digest(w.positions), "workload generator no longer reproduces the fixture's inputs"


def _check_pairs(z, got):
    if "pairs" not in got:
        return
    assert int(z["pair_count"]) == len(got["pairs"])
    assert str(z["pair_sha256"]) == _digest(got["pairs"].astype(np.int32)), "interacting pair set differs from the reference's"


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_port_matches_reference(name):
    z, parts = _load(name)
    w = CASES[name]()
    _check_inputs(z, w)
    for part, flags in PARTS:
        got = evaluate("ReferencePort", w, pairs=part == "direct" and int(z["pair_count"]) >= 0, **flags)
        if part == "direct":
            _check_pairs(z, got)
        assert_parity(parts[part], got, 1e-9, "%s %s [port]" % (name, part))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_b200_matches_reference(name):
    z, parts = _load(name)
    w = CASES[name]()
    _check_inputs(z, w)
    loose = name == "small_nocutoff"     # see test_parity_gpu.test_nonperiodic_methods
    for part, flags in PARTS:
        got = evaluate("B200", w, pairs=part == "direct" and int(z["pair_count"]) >= 0, properties={"Precision": "double"}, **flags)
        if part == "direct":
            _check_pairs(z, got)
        assert_parity(parts[part], got, 1e-5, "%s %s [double]" % (name, part))
        got = evaluate("B200", w, pairs=part == "direct" and int(z["pair_count"]) >= 0, **flags)
        if part == "direct":
            _check_pairs(z, got)
        tol = 1e-3 if loose else (1e-5 if name.startswith(("water", "fep")) else 1e-4)
        if part == "total":
            assert_parity_mixed(parts[part], got, parts["direct"], parts["recip"], what="%s total [mixed]" % name, energy_tol=tol)
        else:
            assert_parity_mixed(parts[part], got, what="%s %s [mixed]" % (name, part), energy_tol=tol)
