#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/check_multi_gpu.py [workload ...]

Every rank takes its share of each evaluation (tile list + exceptions partitioned, reciprocal space on
rank 0 overlapping the reduction, NCCL broadcast / reduce to rank 0); rank 0 -- where the results are
delivered -- then repeats the evaluation alone on a second, single-GPU kernel and compares: forces must
be BIT-IDENTICAL in the deterministic mode (2^32 fixed-point sums do not depend on the partition), slice
energies / derivatives equal to 1e-9 (they travel in 2^32 fixed point), and the partition must cover every
interacting pair exactly once.  Mirrors the reference's testParallelComputation
(platforms/cuda/tests/TestCudaSlicedNonbondedForce.cpp:17-80).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import openmmlab_b200 as mm
    from openmmlab_b200 import systems

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    names = sys.argv[1:] or ["water", "fep", "dhfr"]
    ok = True
    for name in names:
        w = systems.WORKLOADS[name]()
        for share in (None, 1.0, 0.0):
            props = {"DeviceIndex": local, "DeterministicForces": "true"}
            ctx = mm.Context(w.system, "B200", props)
            ctx.setPositions(w.positions if rank == 0 else np.zeros_like(w.positions))    # rank 0's are broadcast
            for k, v in w.parameters.items():
                ctx.setParameter(k, v)
            kernel = ctx._kernels[0][1]
            ident = [kernel.ncclUniqueId() if rank == 0 else None]
            dist.broadcast_object_list(ident, src=0)
            kernel.commInit(ident[0], rank, world)
            if share is not None:
                kernel.setRank0Share(share)
            # 12 evaluations: the first 8 measure and settle the partition, then it is frozen and the evaluation
            # (NCCL all-reduce included) is replayed as a CUDA graph; the LAST one is compared
            for it in range(12):
                ctx._forces = np.zeros((w.num_particles, 3))
                ctx._energyParameterDerivatives = {}
                e = kernel.execute(ctx, True, True, True, True)
            graphs = kernel.getCounters()[4]
            pairs = torch.tensor([kernel.countPairs()])
            dist.all_reduce(pairs)
            if rank == 0:
                solo = mm.Context(w.system, "B200", props)
                solo.setPositions(w.positions)
                for k, v in w.parameters.items():
                    solo.setParameter(k, v)
                k1 = solo._kernels[0][1]
                solo._forces = np.zeros((w.num_particles, 3))
                solo._energyParameterDerivatives = {}
                e1 = k1.execute(solo, True, True, True, True)
                same_forces = np.array_equal(ctx._forces, solo._forces)
                de = abs(e-e1)/max(abs(e1), 1.0)
                ds = np.abs(kernel.sliceEnergies-k1.sliceEnergies).max()/max(np.abs(k1.sliceEnergies).max(), 1.0)
                dd = max(abs(ctx._energyParameterDerivatives[k]-solo._energyParameterDerivatives[k]) for k in solo._energyParameterDerivatives) \
                    if solo._energyParameterDerivatives else 0.0
                same_pairs = int(pairs.item()) == k1.countPairs()
                good = same_forces and (share is not None or graphs == 1) and de < 1e-9 and ds < 1e-9 and dd < 1e-9*max(1.0, abs(e1)) and same_pairs
                ok = ok and good
                print("%-6s world=%d rank0_share=%-5s forces bit-identical=%s  pairs equal=%s  |dE|/E=%.1e  slices=%.1e  derivs=%.1e  graph=%d  %s"
                      % (name, world, share, same_forces, same_pairs, de, ds, dd, graphs, "OK" if good else "FAIL"), flush=True)
            del kernel, ctx
            dist.barrier()
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
