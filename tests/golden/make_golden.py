#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/*.npz.

Run in the BUILD container, where /root/reference exists: the outputs come from platform
"ReferenceRef" = oracle/_ref/libsnb_oracle_ref.so, i.e. the reference's OWN Reference-platform
sources (platforms/reference/src/ReferencePME.cpp, ReferenceSlicedLJCoulombIxn.cpp,
ReferenceSlicedLJCoulomb14.cpp) compiled unchanged.  The fixtures travel to the GPU box, where
/root/reference does not exist; tests/test_golden.py checks the oracle port (CPU) and the CUDA path
(gpu) against them.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_cases import CASES  # noqa: E402
from oracle import oracle_kernel  # noqa: E402
from util import evaluate  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert oracle_kernel.have_ref(), "oracle/_ref is not built (needs /root/reference): make -C oracle ref"
    for name, make in CASES.items():
        w = make()
        has_list = w.force.getNonbondedMethod() != 0      # NoCutoff enumerates all pairs: no list to compare
        out = {}
        for part, flags in (("direct", dict(direct=True, reciprocal=False)), ("recip", dict(direct=False, reciprocal=True)),
                            ("total", dict(direct=True, reciprocal=True))):
            r = evaluate("ReferenceRef", w, pairs=has_list and part == "direct", **flags)
            assert r["kernel"].backend() == "reference"
            out[part+"_energy"] = np.float64(r["energy"])
            out[part+"_slices"] = r["slices"]
            out[part+"_forces"] = r["forces"]
            names = sorted(r["derivs"])
            out[part+"_derivs"] = np.array([r["derivs"][k] for k in names])
            if part == "direct":
                out["pair_count"] = np.int64(len(r["pairs"]) if has_list else -1)
                out["pair_sha256"] = np.array(digest(r["pairs"].astype(np.int32)) if has_list else "")
                out["deriv_names"] = np.array(names)
        out["positions_sha256"] = np.array(digest(w.positions))
        out["num_particles"] = np.int64(w.num_particles)
        np.savez_compressed(os.path.join(HERE, name+".npz"), **out)
        print("%-24s N=%5d pairs=%7d E=%.9g" % (name, w.num_particles, out["pair_count"], out["total_energy"]))


if __name__ == "__main__":
    main()
