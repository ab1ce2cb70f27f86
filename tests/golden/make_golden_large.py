#!/usr/bin/env python
"""Generate the LARGE golden fixtures: the benchmark-size workloads of BASELINE.json (DHFR-, ApoA1- and STMV-size)
evaluated by platform "ReferenceRef" = oracle/_ref, the reference's OWN Reference-platform sources compiled unchanged.

Run in the BUILD container (where /root/reference exists); the STMV-size evaluation takes a few minutes and ~6 GB.
A fixture holds, for the direct-only, reciprocal-only and combined evaluation: total energy, every slice energy,
every derivative, the forces of a strided sample of atoms (every SAMPLE_STRIDE-th atom) with the rms force over ALL
atoms, and for the direct part the number of interacting pairs and the SHA-256 of the sorted (i<j) pair list.
tests/test_golden_large.py checks the CUDA path against them on the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden_large.py [dhfr apoa1 stmv]
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from openmmlab_b200 import systems  # noqa: E402
from oracle import oracle_kernel  # noqa: E402
from util import evaluate  # noqa: E402

SAMPLE_STRIDE = 97


def digest(a):
    h = hashlib.sha256()
    a = np.ascontiguousarray(a)
    flat = a.reshape(-1)
    step = 1 << 24
    for k in range(0, flat.size, step):         # in slices: a 1.3 GB pair list need not be copied whole
        h.update(flat[k:k+step].tobytes())
    return h.hexdigest()


def main():
    assert oracle_kernel.have_ref(), "oracle/_ref is not built (needs /root/reference): make -C oracle ref"
    for name in (sys.argv[1:] or ["dhfr", "apoa1", "stmv"]):
        w = systems.WORKLOADS[name]()
        out = {}
        t0 = time.time()
        for part, flags in (("direct", dict(direct=True, reciprocal=False)), ("recip", dict(direct=False, reciprocal=True)),
                            ("total", dict(direct=True, reciprocal=True))):
            r = evaluate("ReferenceRef", w, pairs=part == "direct", **flags)
            assert r["kernel"].backend() == "reference"
            out[part+"_energy"] = np.float64(r["energy"])
            out[part+"_slices"] = r["slices"]
            out[part+"_forces_sample"] = r["forces"][::SAMPLE_STRIDE].copy()
            out[part+"_rms_force"] = np.float64(np.sqrt((r["forces"]**2).sum(axis=1).mean()))
            names = sorted(r["derivs"])
            out[part+"_derivs"] = np.array([r["derivs"][k] for k in names])
            if part == "direct":
                out["pair_count"] = np.int64(len(r["pairs"]))
                out["pair_sha256"] = np.array(digest(r["pairs"].astype(np.int32, copy=False)))
                out["deriv_names"] = np.array(names)
            del r
        out["positions_sha256"] = np.array(digest(w.positions))
        out["num_particles"] = np.int64(w.num_particles)
        out["sample_stride"] = np.int64(SAMPLE_STRIDE)
        np.savez_compressed(os.path.join(HERE, w.name+"_large.npz"), **out)
        print("%-16s N=%8d pairs=%10d E=%.9g (%.0f s)" % (w.name, w.num_particles, out["pair_count"], out["total_energy"], time.time()-t0), flush=True)


if __name__ == "__main__":
    main()
