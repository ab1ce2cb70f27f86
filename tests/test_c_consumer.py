"""The drop-in boundary used from plain C (tests/c_abi/consumer.c): snb_desc filled by hand, known-answer cases checked against
their closed forms -- independent of the Python mirror (force.py, _abi.PackedDesc) every other test packs its descriptors with.
The same source is built against the CPU oracles (same entry points, prefix snb_oracle_) and against the CUDA library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "consumer.c")


def build_and_run(tmp_path, libdir, libname, oracle):
    exe = str(tmp_path/"consumer")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", libdir, "-l:"+libname, "-Wl,-rpath,"+libdir, "-lm"]
    if oracle:
        cmd.insert(1, "-DSNB_ORACLE")
    else:
        cmd.append("-Wl,-rpath-link,/usr/local/cuda/lib64")
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return subprocess.run([exe], capture_output=True, text=True)


@pytest.mark.parametrize("backend", ["ref", "port"])
def test_c_consumer_against_the_oracle(tmp_path, backend):
    libdir = os.path.join(ROOT, "oracle", "_ref") if backend == "ref" else os.path.join(ROOT, "oracle")
    libname = "libsnb_oracle_ref.so" if backend == "ref" else "libsnb_oracle_port.so"
    if not os.path.exists(os.path.join(libdir, libname)):
        pytest.skip("%s is not built here" % libname)
    out = build_and_run(tmp_path, libdir, libname, oracle=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout+out.stderr


@pytest.mark.gpu
def test_c_consumer_against_the_cuda_library(tmp_path):
    libdir = os.path.join(ROOT, "openmm-lab_b200", "csrc")
    out = build_and_run(tmp_path, libdir, "libslicednb_b200.so", oracle=False)
    assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout+out.stderr


def test_c_consumer_fails_loudly_without_a_gpu(tmp_path):
    """No CPU fallback in the product: without a CUDA device snb_create returns an error and says why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    libdir = os.path.join(ROOT, "openmm-lab_b200", "csrc")
    if not os.path.exists(os.path.join(libdir, "libslicednb_b200.so")):
        pytest.skip("the CUDA library is not built here")
    out = build_and_run(tmp_path, libdir, "libslicednb_b200.so", oracle=False)
    assert out.returncode != 0 and "FAIL create" in out.stdout and "cuda" in out.stdout.lower()
