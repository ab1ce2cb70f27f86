"""Error budget of the single-precision r^2 in the direct-space kernels against the NARROW cutoff band.

The pair set is bit-exact because every pair whose single-precision r^2 falls inside a band around cutoff^2 is decided
again in double, from the untouched coordinates (exactInRange).  That only works if the band is wider than the error
r^2 can carry.  This test replays the kernels' arithmetic in numpy, operation for operation and rounding for rounding
(block-local float coordinates, block centres on the 2^-10 nm lattice, the hi/lo image shift of fixupShift, pairDelta),
on pairs placed at the cutoff -- many of them across the periodic boundary of boxes up to 60 nm -- and checks the
worst error against the band of snb_core.cuh / api.cu: bandTol = 2e-6 rc + 1e-6 rc^2.
"""
import numpy as np
import pytest

f32 = np.float32


def r32(x):
    return np.asarray(x, dtype=np.float64).astype(f32)


def fma32(a, b, c):
    # float32 fma: the product of two float32 is exact in float64; one rounding to float32 at the end
    return r32(a.astype(np.float64)*b.astype(np.float64)+c.astype(np.float64))


def lattice(x):
    return np.rint(x*1024.0)/1024.0


@pytest.mark.parametrize("cutoff", [0.5, 0.9, 1.2, 3.5])
@pytest.mark.parametrize("box", [7.5, 21.6832, 60.0])
def test_r2_error_stays_inside_the_narrow_band(box, cutoff):
    if 0.5*box-cutoff <= 0.35:
        pytest.skip("octets of such a box wrap per pair: wide band")
    rng = np.random.default_rng(int(box*1000+cutoff*10))
    n = 400000
    L = np.array([box, box*0.97, box*1.03])
    # i atom, its block centre (lattice point within 1 nm: SNB_NARROW_MAX_EXTENT) and the octet's centre (block-local)
    xi = rng.random((n, 3))*L
    cI = lattice(xi+rng.uniform(-1.0, 1.0, (n, 3)))
    pi = r32(xi-cI)
    oc = r32(pi.astype(np.float64)+rng.uniform(-0.3, 0.3, (n, 3)))
    # j atom at the cutoff (+- 1e-5), stored wrapped into the box, with its own block centre
    d = rng.normal(size=(n, 3))
    d *= ((cutoff+rng.uniform(-1e-5, 1e-5, n))/np.linalg.norm(d, axis=1))[:, None]
    xj = np.mod(xi+d, L)
    cJ = lattice(xj+rng.uniform(-1.0, 1.0, (n, 3)))
    pj = r32(xj-cJ)
    # fixupShift (direct_impl.cuh): image chosen in single precision, applied as centre difference - m*hi (exact), then m*lo
    hi = r32(lattice(L))
    lo = r32(L-lattice(L))
    inv = r32(1.0/L)
    sx = r32(cJ-cI)
    assert np.array_equal(sx.astype(np.float64), cJ-cI)                 # lattice differences are exact in float
    m = np.rint(r32(r32(r32(pj+sx)-oc)*inv))
    t = fma32(-m, np.broadcast_to(hi, m.shape), sx)
    assert np.array_equal(t.astype(np.float64), (cJ-cI)-m.astype(np.float64)*lattice(L))   # and so is the shift
    q = fma32(-m, np.broadcast_to(lo, m.shape), r32(pj+t))
    # pairDelta + r^2 with the kernels' rounding sequence
    dl = r32(pi-q)
    r2 = fma32(dl[:, 2], dl[:, 2], fma32(dl[:, 1], dl[:, 1], r32(dl[:, 0]*dl[:, 0])))
    # exact: minimum image of the true displacement
    de = xi-xj
    de -= L*np.rint(de/L)
    r2_exact = (de*de).sum(axis=1)
    assert np.allclose(np.sqrt(r2_exact), cutoff, atol=2e-5)
    err = np.abs(r2.astype(np.float64)-r2_exact).max()
    band = min(4e-5*cutoff*cutoff, 2.0e-6*cutoff+1.0e-6*cutoff*cutoff)
    assert err < 0.5*band, "worst |r2 error| %.3e vs band %.3e (box %.1f, cutoff %.2f)" % (err, band, box, cutoff)
