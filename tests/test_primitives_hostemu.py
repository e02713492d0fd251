"""Building blocks of the kernel source (csrc/pam_core.h, compiled for the host by tests/hostemu) against the
third-party routines the reference calls: scipy.optimize.linear_sum_assignment, scipy.ndimage's Gaussian
kernel and 'reflect' boundary, numpy's pairwise summation (np.sum / np.mean), get_believe."""
import ctypes as C

import numpy as np
import pytest
from scipy import ndimage, optimize

from tests.hostemu import build as hb
from tests.util import ptr


@pytest.fixture(scope="module")
def lib():
    L = hb.load()
    L.hostemu_lsap.restype = C.c_int
    L.hostemu_lsap.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.hostemu_gaussian_weights.restype = C.c_int
    L.hostemu_gaussian_weights.argtypes = [C.c_double, C.c_void_p]
    L.hostemu_reflect_index.restype = C.c_int
    L.hostemu_reflect_index.argtypes = [C.c_int, C.c_int]
    L.hostemu_np_sum.restype = C.c_double
    L.hostemu_np_sum.argtypes = [C.c_void_p, C.c_int]
    L.hostemu_mean_confidence.restype = C.c_double
    L.hostemu_mean_confidence.argtypes = [C.c_void_p, C.c_int]
    return L


def _lsap(lib, cost):
    cost = np.ascontiguousarray(cost, np.float64)
    out = np.full(cost.shape[0], -7, np.int32)
    rc = lib.hostemu_lsap(cost.shape[0], cost.shape[1], ptr(cost), ptr(out))
    return rc, out


def test_lsap_equals_scipy_on_random_rectangular_problems(lib):
    """Same assignment as scipy (tracking/IterativeTracker.py:79,150), not just the same cost: continuous random
    costs have a unique optimum."""
    rng = np.random.default_rng(0)
    for _ in range(400):
        nr, nc = int(rng.integers(1, 17)), int(rng.integers(1, 17))
        cost = rng.normal(size=(nr, nc))
        rc, got = _lsap(lib, cost)
        assert rc == 0
        rows, cols = optimize.linear_sum_assignment(cost)
        want = np.full(nr, -1, np.int32)
        want[rows] = cols
        assert np.array_equal(got, want), (nr, nc)


def test_lsap_ties_zero_rows_and_tracker_like_matrices(lib):
    """Degenerate inputs of the tracker: -affinity with many exact zeros (unmatched pairs), duplicate rows
    and columns.  With ties scipy's answer depends on its scan order, which lsap_solve follows."""
    rng = np.random.default_rng(1)
    for _ in range(400):
        nr, nc = int(rng.integers(1, 13)), int(rng.integers(1, 13))
        aff = np.where(rng.random((nr, nc)) < 0.25, rng.uniform(0.1, 1.0, (nr, nc)), 0.0)
        if rng.random() < 0.3 and nr > 1:
            aff[rng.integers(nr)] = aff[rng.integers(nr)]
        if rng.random() < 0.3 and nc > 1:
            aff[:, rng.integers(nc)] = aff[:, rng.integers(nc)]
        cost = -aff
        rc, got = _lsap(lib, cost)
        assert rc == 0
        rows, cols = optimize.linear_sum_assignment(cost)
        want = np.full(nr, -1, np.int32)
        want[rows] = cols
        # the optimum value always agrees; the assignment itself wherever an accepted (positive) pair is concerned
        val = lambda a: sum(cost[i, a[i]] for i in range(nr) if a[i] >= 0)
        assert val(got) == pytest.approx(val(want), abs=1e-12)
        assert sorted(int(c) for c in got if c >= 0) == sorted(set(int(c) for c in got if c >= 0))   # a matching
        assert (got >= 0).sum() == min(nr, nc)
        assert np.array_equal(got, want), (cost, got, want)


def test_lsap_empty_and_infeasible(lib):
    cost = np.full((3, 3), np.inf)
    rc, got = _lsap(lib, cost)
    assert rc == -1                                   # scipy raises "cost matrix is infeasible"
    rc, got = _lsap(lib, np.zeros((0, 4)))
    assert rc == 0


@pytest.mark.parametrize("sigma", [0.3, 0.5, 0.6, 0.8, 1.0, 1.1, 1.4, 2.0])
def test_gaussian_weights_equal_scipy_kernel(lib, sigma):
    """Bit-identical for the sigmas the shipped configurations use (and most others); for a few arguments
    numpy's vectorised exp and libm's exp differ in the last bit (x^2 = 36 at sigma 1.4, x^2 = 1 at sigma 2),
    which moves a weight by one ulp -- neither is the authors' numpy 1.19, so that is the accuracy of the pin."""
    w = np.zeros(9)
    rad = lib.hostemu_gaussian_weights(sigma, ptr(w))
    assert rad == int(4.0 * sigma + 0.5)
    ref = ndimage._filters._gaussian_kernel1d(sigma, 0, rad)[rad:]
    if sigma <= 1.1:
        assert np.array_equal(w[: rad + 1], ref)
    else:
        assert np.all(np.abs(w[: rad + 1] - ref) <= 2.5e-16 * ref)


def test_gaussian_radius_limit(lib):
    w = np.zeros(9)
    assert lib.hostemu_gaussian_weights(2.2, ptr(w)) == -1     # radius 9 > 8: rejected by pam_create


def test_reflect_index_equals_scipy_reflect_mode(lib):
    """Index map of mode='reflect' (d c b a | a b c d | d c b a), read off scipy itself."""
    for n in range(1, 13):
        x = np.arange(n, dtype=np.float64)
        for rad in range(0, 9):
            # correlate with a one-hot kernel picks x[reflect(i + off)]
            for off in (-rad, rad):
                k = np.zeros(2 * rad + 1)
                k[rad + off] = 1.0
                picked = ndimage.correlate1d(x, k, mode="reflect")
                for i in (0, n - 1):
                    assert lib.hostemu_reflect_index(i + off, n) == int(picked[i]), (n, rad, off, i)


def test_np_sum_is_numpy_pairwise_sum(lib):
    rng = np.random.default_rng(3)
    for n in list(range(0, 40)) + [63, 64, 65, 100, 127, 128]:      # n <= 128: the leaf of numpy's pairwise sum (all uses: J <= 32)
        x = (rng.normal(size=n) * 10.0 ** rng.uniform(-3, 3, n)).astype(np.float64)
        assert lib.hostemu_np_sum(ptr(x), n) == float(np.sum(x)), n


def test_mean_confidence_is_get_believe(lib):
    rng = np.random.default_rng(4)
    for J in (5, 8, 14, 17, 19, 23, 32):
        for neg in (0, 1, 3, J):
            pose = rng.uniform(0, 1, (J, 3)).astype(np.float32)
            if neg:
                pose[rng.choice(J, neg, replace=False), 2] = -1.0
            got = lib.hostemu_mean_confidence(ptr(pose), J)
            sel = [np.float64(p[2]) for p in pose.astype(np.float64) if p[2] >= 0]
            if not sel:
                assert np.isnan(got)
            else:
                assert got == float(np.mean(sel)), (J, neg)


def test_small_assignment_solvers_agree_with_scipy(lib):
    """The kernel's closed-form and enumerated assignment (csrc/pam_track.h: assign_closed_form, assign_enumerated) on
    random sparse affinity matrices: the matched (track, detection) pairs with positive affinity are those of
    scipy.optimize.linear_sum_assignment(-A) (IterativeTracker.py:150-160).  Exact ties between different matchings
    (duplicated columns / rows) must be handed to the general solver, which follows scipy's order."""
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(5)
    stages = {0: 0, 1: 0, 2: 0, 3: 0}

    def check(A, limit=1024):
        n, mm = A.shape
        t2d = np.zeros(n, np.int32)
        A = np.ascontiguousarray(A, dtype=np.float64)
        how = lib.hostemu_assign(A.ctypes.data_as(C.c_void_p), n, mm, limit, t2d.ctypes.data_as(C.c_void_p))
        stages[how] += 1
        rows, cols = linear_sum_assignment(-A)
        want = np.full(n, -1, np.int32)
        for r, c in zip(rows, cols):
            if A[r, c] > 0:
                want[r] = c
        assert np.array_equal(t2d, want), (how, A, t2d, want)
        return how

    for trial in range(4000):
        n, mm = int(rng.integers(1, 9)), int(rng.integers(1, 7))
        A = rng.uniform(0.05, 1.0, (n, mm)) * (rng.random((n, mm)) < rng.choice([0.15, 0.3, 0.6]))
        check(A)
    for trial in range(300):                                   # larger problems: enumeration limit, then the general solver
        n, mm = int(rng.integers(6, 20)), int(rng.integers(5, 14))
        A = rng.uniform(0.05, 1.0, (n, mm)) * (rng.random((n, mm)) < 0.5)
        check(A, limit=int(rng.choice([32, 1024, 4096])))
    # structural ties: two identical detections / two identical tracks -> general solver (scipy's pivoting order)
    for trial in range(300):
        n, mm = int(rng.integers(2, 7)), int(rng.integers(2, 6))
        A = rng.uniform(0.05, 1.0, (n, mm)) * (rng.random((n, mm)) < 0.7)
        if trial % 2:
            A[:, -1] = A[:, 0]
        else:
            A[-1, :] = A[0, :]
        check(A)
    assert all(v > 0 for v in stages.values()), stages
    print("assignment stages used (quick, closed form, enumeration, general):", stages)
