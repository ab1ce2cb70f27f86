"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): the partitioned evaluation over
NCCL must reproduce the single-GPU one -- forces bit for bit, energies to 1e-12 (tools/check_multi_gpu.py).
Mirrors the reference's testParallelComputation (platforms/cuda/tests/TestCudaSlicedNonbondedForce.cpp:17-80)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_partitioned_evaluation_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tools", "check_multi_gpu.py"), "water", "fep"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:]+out.stderr[-3000:]
