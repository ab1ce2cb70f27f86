import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import openmmlab_b200 as mm
from openmmlab_b200 import systems, SlicedNonbondedForce as F
from oracle import oracle_kernel
from util import evaluate, max_force_error

def report(w, **kw):
    ref = evaluate("Reference", w, pairs=w.force.getNonbondedMethod()!=0, **kw)
    got = evaluate("B200", w, pairs=True, **kw)
    print("==", w.name, w.force.getNonbondedMethodName(), kw)
    if "pairs" in ref:
        a = set(map(tuple, ref["pairs"])); b = set(map(tuple, got["pairs"]))
        print(" pairs ref %d got %d (raw %d) missing %d extra %d" % (len(a), len(b), len(got["pairs"]), len(a-b), len(b-a)))
        if a-b: print("  missing sample", sorted(a-b)[:8])
        if b-a: print("  extra sample", sorted(b-a)[:8])
    print(" E ref %.6f got %.6f" % (ref["energy"], got["energy"]))
    print(" slices ref", ref["slices"].ravel()); print(" slices got", got["slices"].ravel())
    print(" derivs", ref["derivs"], got["derivs"])
    print(" max force err %.3e" % max_force_error(ref["forces"], got["forces"]), "nan forces:", np.isnan(got["forces"]).sum())
    k = got["kernel"]; print(" counters", k.getCounters())

# 5 atom chain cutoff nonperiodic
system = mm.System(); f = F(1); f.setNonbondedMethod(F.CutoffNonPeriodic)
for i in range(5): system.addParticle(1); f.addParticle(0, 1.5, 1 if i in (0,1) else 0)
f.setCutoffDistance(3.5); f.createExceptionsFromBonds([(0,1),(1,2),(2,3),(3,4)], 0, 0)
system.addForce(f)
w = systems.Workload("chain5", system, f, np.array([[i,0,0] for i in range(5)], float), None, {}, {})
report(w)
report(systems.water_box(method=F.CutoffPeriodic))
report(systems.water_box(method=F.PME), reciprocal=False)
report(systems.water_box(method=F.PME), direct=False)
report(systems.water_box(method=F.PME))
