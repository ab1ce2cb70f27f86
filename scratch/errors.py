import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import openmmlab_b200 as mm
from openmmlab_b200 import systems, SlicedNonbondedForce as F
from oracle import oracle_kernel
from util import evaluate, max_force_error, rel_err

def report(w, **kw):
    ref = evaluate("Reference", w, **kw)
    got = evaluate("B200", w, **kw)
    print("==", w.name, w.force.getNonbondedMethodName(), kw)
    print("  E ref %.6f got %.6f abs %.2e rel %.2e" % (ref["energy"], got["energy"], abs(ref["energy"]-got["energy"]), rel_err(ref["energy"], got["energy"])))
    for s in range(ref["slices"].shape[0]):
        for t in range(2):
            a, b = ref["slices"][s,t], got["slices"][s,t]
            print("  slice %d %s ref %14.6f abs %.2e rel %.2e" % (s, "CV"[t], a, abs(a-b), rel_err(a,b)))
    fe = np.abs(ref["forces"]-got["forces"]).max(axis=1)
    scale = np.maximum(np.linalg.norm(ref["forces"],axis=1),1)
    i = np.argmax(fe/scale)
    print("  force max rel %.2e (atom %d |F|=%.1f abs %.2e)  max abs %.2e  rms abs %.2e" % ((fe/scale).max(), i, scale[i], fe[i], fe.max(), np.sqrt(((ref["forces"]-got["forces"])**2).mean())))

w = systems.water_box()
report(w, reciprocal=False); report(w, direct=False); report(w)
w = systems.small_mixed(method=F.PME, num_groups=2, scaling=[("lam_a", 0.3, 0, 1, True, True)], derivatives=["lam_a"])
report(w, reciprocal=False); report(w, direct=False)
report(systems.fep_ligand(), reciprocal=False); report(systems.fep_ligand(), direct=False)
w = systems.dhfr_like()
report(w, reciprocal=False); report(w, direct=False); report(w)
