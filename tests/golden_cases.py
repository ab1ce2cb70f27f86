"""The workloads behind tests/golden/*.npz (shared by the generator and the tests)."""
import numpy as np

from openmmlab_b200 import SlicedNonbondedForce as F
from openmmlab_b200 import systems


def _triclinic():
    w = systems.small_mixed(method=F.PME, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"])
    L = w.box[0][0]
    box = np.array([[L, 0, 0], [0.25*L, L, 0], [-0.15*L, 0.3*L, L]]).astype(np.float32).astype(np.float64)
    w.box = box
    w.system.setDefaultPeriodicBoxVectors(*box)
    return w


_SCALING3 = [("lam_a", 0.3, 0, 1, True, True), ("lam_b", 0.65, 1, 1, True, False), ("lam_c", 0.9, 0, 2, False, True)]

CASES = {
    # BASELINE.json configs[0]
    "water648_cutoff": lambda: systems.water_box(method=F.CutoffPeriodic),
    "water648_pme": lambda: systems.water_box(method=F.PME),
    "small_pme_3groups": lambda: systems.small_mixed(method=F.PME, num_groups=3, scaling=_SCALING3, derivatives=["lam_a", "lam_b"]),
    "small_ljpme_2groups": lambda: systems.small_mixed(method=F.LJPME, num_groups=2, scaling=_SCALING3[:2], derivatives=["lam_a", "lam_b"],
                                                       ljpme=(systems.ALPHA, 12, 12, 12)),
    "small_pme_switch": lambda: systems.small_mixed(method=F.PME, switch=0.75, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"]),
    "small_pme_triclinic": _triclinic,
    "small_nocutoff": lambda: systems.small_mixed(method=F.NoCutoff, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"]),
    # BASELINE.json configs[3]
    "fep3000_ljpme": systems.fep_ligand,
}
