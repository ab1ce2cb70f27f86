import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import openmmlab_b200 as mm
from openmmlab_b200 import systems, SlicedNonbondedForce as F
from oracle import oracle_kernel
from util import evaluate, max_force_error
system = mm.System(); f = F(1); f.setNonbondedMethod(F.CutoffNonPeriodic)
for i in range(5): system.addParticle(1); f.addParticle(0, 1.5, 1 if i in (0,1) else 0)
f.setCutoffDistance(3.5); f.createExceptionsFromBonds([(0,1),(1,2),(2,3),(3,4)], 0, 0)
system.addForce(f)
w = systems.Workload("chain5", system, f, np.array([[i,0,0] for i in range(5)], float), None, {}, {})
got = evaluate("B200", w, pairs=True)
print(got["pairs"][:10], got["energy"])
lib = mm.load_library()
lib.snb_debug_dump.argtypes=[ctypes.c_void_p]
lib.snb_debug_dump(got["kernel"]._handle)
