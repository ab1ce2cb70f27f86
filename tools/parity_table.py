#!/usr/bin/env python
"""Residual table of the CUDA path against the reference's own sources: the five BASELINE.json configurations x
{mixed with the general direct-space kernel, mixed with the packed one, Precision=double} x {direct, reciprocal, total},
by the reference's metric (openmmapi/include/internal/AssertionUtilities.h:7-37).

    python tools/parity_table.py [workload ...] > profiles/rN_parity_table.json      (markdown on stderr)

Reference side: oracle/_ref run live for the small workloads (water, fep); the committed fixtures generated from oracle/_ref
in the build container for the benchmark-size ones (tests/golden/*_large.npz: strided force sample, every 97th atom).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PARTS = (("direct", dict(direct=True, reciprocal=False)), ("recip", dict(direct=False, reciprocal=True)),
         ("total", dict(direct=True, reciprocal=True)))


def main():
    import bench
    from openmmlab_b200 import systems
    from util import evaluate
    names = sys.argv[1:] or ["water", "dhfr", "apoa1", "fep", "stmv"]
    table = []
    for name in names:
        w = systems.WORKLOADS[name]()
        fixture = os.path.join(ROOT, "tests", "golden", w.name+"_large.npz")
        refs, stride = {}, 1
        if os.path.exists(fixture):
            z = np.load(fixture)
            dn = [str(s) for s in z["deriv_names"]]
            stride = int(z["sample_stride"])
            for part, _ in PARTS:
                refs[part] = {"energy": float(z[part+"_energy"]), "slices": z[part+"_slices"], "forces": z[part+"_forces_sample"],
                              "derivs": dict(zip(dn, z[part+"_derivs"])), "rms_force": float(z[part+"_rms_force"])}
            pairs_ref = (int(z["pair_count"]), str(z["pair_sha256"]))
            source = "fixture from oracle/_ref"
        else:
            for part, flags in PARTS:
                r = evaluate("Reference", w, pairs=part == "direct", **flags)
                refs[part] = r
                if part == "direct":
                    pairs_ref = (len(r["pairs"]), bench.digest(r["pairs"].astype(np.int32)))
            source = "oracle (%s) run live" % refs["total"]["kernel"].backend()
        packed_ok = w.force.getNumSubsets() <= 2 and w.force.getNonbondedMethodName() == "PME"
        modes = [("mixed/general", {"DirectKernel": "general"})]+([("mixed/packed", {"DirectKernel": "packed"})] if packed_ok else []) + \
                [("double", {"Precision": "double"})]
        for mode, props in modes:
            for part, flags in PARTS:
                if part == "recip" and mode == "mixed/packed":
                    continue
                got = evaluate("B200", w, pairs=part == "direct", properties=props, **flags)
                r = bench.residuals(refs[part], got, stride)
                row = {"workload": w.name, "atoms": w.num_particles, "mode": mode, "part": part, "reference": source}
                row.update(r)
                if part == "direct":
                    row["pair_count"] = int(len(got["pairs"]))
                    row["pair_set_equal"] = bool((len(got["pairs"]), bench.digest(got["pairs"].astype(np.int32))) == pairs_ref)
                row["within_1e-5"] = bool(max(r["energy_rel"], r["worst_slice_rel"], r["worst_derivative_rel"], r["force_per_atom_metric"]) <= 1e-5)
                table.append(row)
                print("| %-12s | %-13s | %-6s | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e | %s | %s |" % (
                    w.name, mode, part, r["energy_rel"], r["worst_slice_rel"], r["worst_derivative_rel"], r["force_per_atom_metric"],
                    r["force_rms_over_rms_force"], r["force_max_over_rms_force"], row.get("pair_set_equal", ""), row["within_1e-5"]), file=sys.stderr, flush=True)
                print("    worst slice: %r" % (r["worst_slice"],), file=sys.stderr, flush=True)
                del got
    print(json.dumps(table))


if __name__ == "__main__":
    main()
