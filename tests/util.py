"""Shared helpers for the test-suite: the reference's own comparison metric and the
platform matrix (CPU oracle platforms unmarked, CUDA platform marked gpu)."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle_kernel  # noqa: E402

ONE_4PI_EPS0 = 138.93545764438198   # include/snb_constants.h

CPU_PLATFORMS = (["ReferenceRef"] if oracle_kernel.have_ref() else [])+["ReferencePort"]
PLATFORMS = CPU_PLATFORMS+[pytest.param("B200", marks=pytest.mark.gpu)]


def assert_equal_to(expected, found, tol):
    # assertEqualTo, openmmapi/include/internal/AssertionUtilities.h:7-14
    scale = max(abs(expected), 1.0)
    assert abs(expected-found)/scale <= tol, "Expected %r, found %r (rel %g > %g)" % (expected, found, abs(expected-found)/scale, tol)


def assert_equal_vec(expected, found, tol):
    # assertEqualVec, AssertionUtilities.h:16-27
    expected = np.asarray(expected, dtype=float)
    found = np.asarray(found, dtype=float)
    scale = max(math.sqrt(expected.dot(expected)), 1.0)
    err = np.abs(expected-found).max()/scale
    assert err <= tol, "Expected %r, found %r (rel %g > %g)" % (expected, found, err, tol)


def max_force_error(expected, found):
    """max over atoms/components of |dF| / max(|F_ref|, 1)  (assertForces metric)."""
    expected = np.asarray(expected, dtype=float)
    found = np.asarray(found, dtype=float)
    scale = np.maximum(np.sqrt((expected**2).sum(axis=1)), 1.0)
    return (np.abs(expected-found).max(axis=1)/scale).max()


def rel_err(expected, found):
    return abs(expected-found)/max(abs(expected), 1.0)
