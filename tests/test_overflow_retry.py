"""Tile-list overflow: an evaluation whose list outgrows its buffers is discarded on the device, the host grows the
buffers and repeats.  Nothing of the discarded pass may reach a caller-owned accumulator -- the device entry point
ADDS its forces into the caller's buffer (include/slicednb.h, snb_execute_device), so a pass counted twice would
double them silently.  The capacity is forced down through the test hook snb_set_list_capacity."""
import numpy as np
import pytest

import openmmlab_b200 as mm
from openmmlab_b200 import _abi, systems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("direct_kernel", ["general", "packed"])
def test_device_forces_survive_a_list_overflow(direct_kernel):
    import torch
    w = systems.small_mixed(n_side=8, solute_sites=20, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"])
    n = w.num_particles
    # deterministic mode: forces are bit-reproducible from run to run, so "equal" can mean equal
    ctx = mm.Context(w.system, "B200", {"DirectKernel": direct_kernel, "DeterministicForces": "true"})
    ctx.setPositions(w.positions)
    for k, v in w.parameters.items():
        ctx.setParameter(k, v)
    kernel = ctx._kernels[0][1]
    # the host path as the yardstick (its result buffer is overwritten by every pass, so a repeat cannot hurt it)
    ctx._forces = np.zeros((n, 3))
    ctx._energyParameterDerivatives = {}
    e_host = kernel.execute(ctx, True, True, True, True)
    chunks_needed = kernel.getCounters()[1]
    assert chunks_needed > 64
    posq = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    posq[:, :3] = torch.from_numpy(w.positions.astype(np.float32)).cuda()
    box = np.ascontiguousarray(w.box, dtype=np.float64).reshape(9)
    params = np.array([ctx.getParameter(name) for name in kernel.globalNames], dtype=np.float64)

    def device_forces():
        acc = torch.zeros((3, n), dtype=torch.int64, device="cuda")
        kernel.executeDevice(posq.data_ptr(), box, params, _abi.EXECUTE_ALL, acc.data_ptr())
        torch.cuda.synchronize()
        return acc.cpu().numpy().T/2.0**32, kernel.readEnergies(_abi.EXECUTE_ALL)[0]

    f_ok, e_ok = device_forces()
    # far too small: the first pass overflows, the buffers grow, the evaluation repeats
    kernel.setListCapacity(max(8, chunks_needed//5))
    f_retry, e_retry = device_forces()
    assert kernel.getCounters()[1] == chunks_needed
    assert np.array_equal(f_retry, f_ok), "a discarded pass leaked into the caller's force accumulator (max |dF| = %g)" % np.abs(f_retry-f_ok).max()
    assert abs(e_retry-e_ok) <= 1e-12*max(1.0, abs(e_ok))     # (slice energies are summed with FP64 atomics: order-dependent last bits)
    # and both agree with the host path (same fixed-point sums, converted the same way)
    assert np.abs(f_ok-ctx._forces).max() <= 1e-9*max(1.0, np.abs(ctx._forces).max())
    assert abs(e_ok-e_host) <= 1e-9*max(1.0, abs(e_host))
