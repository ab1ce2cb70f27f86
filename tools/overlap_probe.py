#!/usr/bin/env python
"""How well do the two chains of an evaluation overlap?  Device time (CUDA events around the replayed graph) of the
direct-space part alone, the reciprocal-space part alone and the whole evaluation.

    python tools/overlap_probe.py [dhfr|fep|apoa1|stmv] [steps]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import openmmlab_b200 as mm  # noqa: E402
from openmmlab_b200 import _abi, systems  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "dhfr"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    w = systems.WORKLOADS[name]()
    n = w.num_particles
    ctx = mm.Context(w.system, "B200", {})
    ctx.setPositions(w.positions)
    ctx.setPeriodicBoxVectors(*w.box)
    for k, v in w.parameters.items():
        ctx.setParameter(k, v)
    kernel = ctx._kernels[0][1]
    posq = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    posq[:, :3] = torch.from_numpy(w.positions.astype(np.float32)).cuda()
    box = np.ascontiguousarray(w.box, dtype=np.float64).reshape(9)
    params = np.array([ctx.getParameter(p) for p in kernel.globalNames], dtype=np.float64)
    acc = torch.zeros((3, n), dtype=torch.int64, device="cuda")
    flush = torch.empty(64*1024*1024, dtype=torch.float32, device="cuda")
    out = {"workload": name, "atoms": n}
    F, E, D, R = _abi.INCLUDE_FORCES, _abi.INCLUDE_ENERGY, _abi.INCLUDE_DIRECT, _abi.INCLUDE_RECIPROCAL
    for label, flags in (("direct", F | E | D), ("reciprocal", F | E | R), ("both", F | E | D | R)):
        ms = []
        for it in range(steps+5):
            flush.fill_(0.0)
            torch.cuda.synchronize()
            kernel.executeDevice(posq.data_ptr(), box, params, flags, acc.data_ptr())
            if it >= 5:
                ms.append(kernel.getLastDeviceMs())
        out[label+"_us"] = round(1e3*float(np.median(ms)), 2)
    out["sum_us"] = round(out["direct_us"]+out["reciprocal_us"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
