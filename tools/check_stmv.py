#!/usr/bin/env python
"""Full-size parity of the STMV-size workload (BASELINE.json configs[4], 1,066,628 atoms, PME 196^3) against the
CPU oracle: one evaluation each (the oracle needs about a minute), interacting-pair COUNT equal (the sorted lists
themselves are compared at smaller sizes by the test-suite), energies / derivatives / forces by the mixed criteria
of tests/util.py.  Prints one JSON line; exit code 1 on failure.

    python tools/check_stmv.py [> profiles/rN_stmv_parity.json]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from openmmlab_b200 import systems
    from util import assert_parity_mixed, evaluate, max_force_error, rel_err
    w = systems.stmv_like()
    t0 = time.time()
    got = evaluate("B200", w)
    pairs_gpu = got["kernel"].countPairs()
    t_gpu = time.time()-t0
    # the oracle's direct and reciprocal parts on their own (they cancel to ~1e-3 of their size in the total:
    # util.assert_parity_mixed scales the total's tolerance by them), checked on their own as well
    refs, failures = {}, []
    t0 = time.time()
    for part, flags in (("direct", dict(direct=True, reciprocal=False)), ("recip", dict(direct=False, reciprocal=True))):
        refs[part] = evaluate("Reference", w, **flags)
        try:
            assert_parity_mixed(refs[part], evaluate("B200", w, **flags), what="stmv %s [mixed]" % part, energy_tol=1e-5)
        except AssertionError as e:
            failures.append(str(e))
    ref = evaluate("Reference", w)
    t_ref = time.time()-t0
    # the oracle's neighbour list of that evaluation: its size (oracle/harness.cpp snb_oracle_count_pairs)
    import ctypes
    count = ref["kernel"]._lib.snb_oracle_count_pairs
    count.restype, count.argtypes = ctypes.c_longlong, [ctypes.c_void_p]
    pairs_ref = int(count(ref["kernel"]._handle))
    dF = ref["forces"]-got["forces"]
    rmsF = np.sqrt((ref["forces"]**2).sum(axis=1).mean())
    gross = np.abs(ref["slices"]).sum()
    out = {"workload": w.name, "atoms": w.num_particles, "oracle": ref["kernel"].backend(), "oracle_seconds_direct_recip_total": t_ref, "gpu_seconds_first_call": t_gpu,
           "pairs_gpu": int(pairs_gpu), "pairs_oracle": pairs_ref,
           "energy_ref": ref["energy"], "energy_gpu": got["energy"], "energy_rel_err": rel_err(ref["energy"], got["energy"]),
           "slice_abs_err_over_gross": float(np.abs(ref["slices"]-got["slices"]).max()/gross),
           "deriv_rel_err": max(rel_err(ref["derivs"][k], got["derivs"][k]) for k in ref["derivs"]),
           "force_rms_err_over_rms_force": float(np.sqrt((dF**2).sum(axis=1).mean())/rmsF),
           "force_max_err_over_rms_force": float(np.abs(dF).max()/rmsF),
           "force_per_atom_reference_metric": float(max_force_error(ref["forces"], got["forces"]))}
    try:
        assert_parity_mixed(ref, got, refs["direct"], refs["recip"], what="stmv total [mixed]", energy_tol=1e-5)
    except AssertionError as e:
        failures.append(str(e))
    ok = pairs_ref == pairs_gpu and not failures
    out["failures"] = failures
    out["ok"] = bool(ok)
    print(json.dumps(out))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
