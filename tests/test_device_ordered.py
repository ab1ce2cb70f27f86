"""snb_execute_device_ordered: the zero-copy entry point for a binding to OpenMM's CUDA platform (SURVEY.md 8f rank 1;
plugin/src/B200CudaOpenMMLabKernels.cpp).  Positions and forces live in the platform's own REORDERED, padded atom order:
posq float4 (+ float4 correction in mixed precision) or double4, the atom-index array, a [3][padded] 2^32 fixed-point
force buffer that is ADDED to (reference: platforms/cuda/src/CudaOpenMMLabKernels.cpp:736-758, force layout
platforms/common/src/kernels/pme.cc:399-405).  Checked against the host entry point on the same positions."""
import numpy as np
import pytest

import openmmlab_b200 as mm
from openmmlab_b200 import _abi, systems

pytestmark = pytest.mark.gpu


def _host(w, props):
    ctx = mm.Context(w.system, "B200", props)
    ctx.setPositions(w.positions)
    for k, v in w.parameters.items():
        ctx.setParameter(k, v)
    kernel = ctx._kernels[0][1]
    ctx._forces = np.zeros((w.num_particles, 3))
    ctx._energyParameterDerivatives = {}
    e = kernel.execute(ctx, True, True, True, True)
    return ctx, kernel, e, ctx._forces.copy()


@pytest.mark.parametrize("mode", ["single", "mixed", "double"])
def test_reordered_padded_layout(mode):
    import torch
    w = systems.small_mixed(n_side=8, solute_sites=20, scaling=[("lam", 0.5, 0, 1, True, True)], derivatives=["lam"])
    n = w.num_particles
    props = {"DeterministicForces": "true"} if mode != "double" else {"Precision": "double"}
    ctx, kernel, e_host, f_host = _host(w, props)
    slices_host = kernel.sliceEnergies.copy()
    rng = np.random.default_rng(3)
    order = rng.permutation(n).astype(np.int32)             # slot t holds atom order[t]
    padded = 32*((n+31)//32)+32
    pos = w.positions[order]
    dev = "cuda"
    if mode == "double":
        posq = torch.zeros((padded, 4), dtype=torch.float64, device=dev)
        posq[:n, :3] = torch.from_numpy(pos).to(dev)
        corr = None
    else:
        p32 = pos.astype(np.float32)
        posq = torch.zeros((padded, 4), dtype=torch.float32, device=dev)
        posq[:n, :3] = torch.from_numpy(p32).to(dev)
        corr = None
        if mode == "mixed":
            corr = torch.zeros((padded, 4), dtype=torch.float32, device=dev)
            corr[:n, :3] = torch.from_numpy((pos-p32.astype(np.float64)).astype(np.float32)).to(dev)
    index = torch.from_numpy(order).to(dev)
    # the platform's force buffer already holds other forces: ours are ADDED
    before = rng.integers(-2**40, 2**40, size=(3, padded), dtype=np.int64)
    forces = torch.from_numpy(before.copy()).to(dev)
    box = np.ascontiguousarray(w.box, dtype=np.float64).reshape(9)
    params = np.array([ctx.getParameter(name) for name in kernel.globalNames], dtype=np.float64)
    for rep in range(3):        # eager, captured, replayed: the graph must carry the layout
        forces.copy_(torch.from_numpy(before).to(dev))
        kernel.executeDeviceOrdered(posq.data_ptr(), box, params, _abi.EXECUTE_ALL, forces.data_ptr(), padded, index.data_ptr(),
                                    corr.data_ptr() if corr is not None else None, mode == "double")
        torch.cuda.synchronize()
        got = forces.cpu().numpy()
        e, slices, derivs = kernel.readEnergies(_abi.EXECUTE_ALL)
        assert np.array_equal(got[:, n:], before[:, n:]), "padding slots of the force buffer were touched"
        f = np.empty((n, 3))
        f[order] = ((got-before)[:, :n].T)/2.0**32
        rms = np.sqrt((f_host**2).sum(axis=1).mean())
        # single: the positions themselves were rounded to float (3e-7 nm at 3 nm): forces move by ~1e-4 of their size;
        # mixed: posq + correction is the double position to ~1e-14 nm, which can still flip the float rounding of a block-local
        # coordinate (FP32 pair arithmetic); double: the same numbers in, the same numbers out
        tol = {"single": 2e-3, "mixed": 1e-5, "double": 1e-9}[mode]
        assert np.abs(f-f_host).max() <= tol*rms, "%s rep %d: forces differ by %g of the rms force" % (mode, rep, np.abs(f-f_host).max()/rms)
        etol = {"single": 1e-5, "mixed": 1e-7, "double": 1e-10}[mode]*max(1.0, np.abs(slices_host).sum())
        assert abs(e-e_host) <= etol and np.abs(slices-slices_host).max() <= etol


def test_bad_layout_is_refused():
    import torch
    w = systems.water_box()
    ctx, kernel, _, _ = _host(w, {})
    posq = torch.zeros((w.num_particles, 4), dtype=torch.float32, device="cuda")
    box = np.ascontiguousarray(w.box, dtype=np.float64).reshape(9)
    params = np.array([ctx.getParameter(name) for name in kernel.globalNames], dtype=np.float64)
    with pytest.raises(mm.SnbError):
        kernel.executeDeviceOrdered(posq.data_ptr(), box, params, _abi.EXECUTE_ALL, None, w.num_particles-1)
