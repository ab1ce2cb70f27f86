import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from openmmlab_b200 import systems
from util import evaluate, assert_parity_mixed
name = sys.argv[1] if len(sys.argv) > 1 else "water"
w = systems.WORKLOADS[name]()
ref = evaluate("Reference", w)
got = evaluate("B200", w)
print("E ref %.6f got %.6f" % (ref["energy"], got["energy"]))
print("max |dF| %.3e  rms F %.3e" % (np.abs(ref["forces"]-got["forces"]).max(), np.sqrt((ref["forces"]**2).sum(1).mean())))
print("slices ref", ref["slices"].ravel(), "\nslices got", got["slices"].ravel())
