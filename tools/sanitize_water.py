#!/usr/bin/env python
"""The workload of the compute-sanitizer runs: the 648-atom water box, a few evaluations of every
phase -- PME and CutoffPeriodic, both direct-space kernels, mixed and double precision, deterministic spreading.

    compute-sanitizer --tool memcheck|racecheck python tools/sanitize_water.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import openmmlab_b200 as mm
    from openmmlab_b200 import SlicedNonbondedForce as F
    from openmmlab_b200 import systems
    for method in (F.PME, F.CutoffPeriodic):
        w = systems.water_box(method=method)
        for props in ({}, {"DirectKernel": "packed"}, {"Precision": "double"}, {"DeterministicForces": "true"}, {"ListSkin": "0.02"}):
            ctx = mm.Context(w.system, "B200", props)
            ctx.setPositions(w.positions)
            for k, v in w.parameters.items():
                ctx.setParameter(k, v)
            kernel = ctx._kernels[0][1]
            for it in range(3):      # eager, captured, replayed
                ctx._forces = np.zeros((w.num_particles, 3))
                ctx._energyParameterDerivatives = {}
                e = kernel.execute(ctx, True, True, True, True)
            print(w.name, method, props, "E = %.6f" % e, flush=True)
            del kernel, ctx


if __name__ == "__main__":
    main()
