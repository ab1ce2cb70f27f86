"""The reference's OWN Python tests of SlicedNonbondedForce (python/tests/TestSlicedNonbondedForce.py), run UNCHANGED from where
they lie in /root/reference on this repository's host mirror: `import openmm` / `import openmmlab` resolve to the stand-ins of
tests/ref_python_shim/.  The "Reference" ids run on the CPU oracle here (the "CUDA" ids would run the B200 kernel; the
reference's sources are not on the GPU box, so they are not part of the GPU suite)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/python/tests/TestSlicedNonbondedForce.py"


def run_reference_tests(ids):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests", "ref_python_shim")+os.pathsep+ROOT)
    # (node ids, not -k: -k also matches the directory names of the path; -p no:cacheprovider: the reference tree is read-only;
    # --rootdir / -c keep this repository's conftest and ini out of the run)
    nodes = [REF_TESTS+"::"+i for i in ids]
    return subprocess.run([sys.executable, "-m", "pytest"]+nodes+["-q", "-p", "no:cacheprovider", "--rootdir", os.path.dirname(REF_TESTS),
                           "-c", os.devnull], capture_output=True, text=True, env=env, cwd="/tmp")


@pytest.mark.skipif(not os.path.isfile(REF_TESTS), reason="the reference's sources are not on this machine")
def test_reference_python_tests_on_the_oracle():
    out = run_reference_tests(["testParameterClash[Reference]", "testCoulomb[Reference]", "testLargeSystem[Reference]"])
    assert out.returncode == 0, out.stdout[-3000:]+out.stderr[-2000:]
    assert "3 passed" in out.stdout, out.stdout[-1000:]
