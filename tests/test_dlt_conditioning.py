"""Where parity of the triangulation is meaningful -- and where it is not.

The reference weights a view that is T frames old by exp(-lambda_t T).  With the shipped
lambda_t = 5 the smallest weight is e^-15 = 3e-7 and LAPACK's SVD (absolute accuracy eps * sigma_max)
still resolves the solution to ~1e-8 relative: the CUDA path agrees with it far inside the 0.5 mm bar
(test_dlt_matches_lapack_for_shipped_weights).  With e.g. lambda_t = 10 a 3-frame-old view gets the weight
e^-30 = 9e-14: a system that rests on such rows has a condition number > 1e13 and LAPACK's answer carries
~1e-3 relative rounding noise (millimetres), which no other implementation can reproduce.  There the
Givens + one-sided-Jacobi path of this repo, which is accurate relative to each row's own scale, returns
a smallest singular vector to working precision that differs from LAPACK's by no more than the
conditioning allows (test_ill_conditioned_system_is_solved_to_working_precision)."""
import ctypes as C

import numpy as np

from tests.hostemu import build as hb
from tests.util import ptr
from pam_b200 import camera, synth


def _systems(rng, P, n, ages, lam, noise=1.0):
    V = len(P)
    X = np.c_[rng.uniform(-1.5, 1.5, (n, 2)), rng.uniform(0, 1.8, n)]
    pr = np.einsum("vik,nk->nvi", P, np.c_[X, np.ones(n)])
    uv = pr[..., :2] / pr[..., 2:] + rng.normal(0, noise, (n, V, 2))
    w = np.exp(-lam * ages)
    return X, uv, w


def _rows(P, uv, w):
    rows = []
    for a in range(len(P)):
        u, v = uv[a]
        for r in (u * P[a, 2] - P[a, 0], v * P[a, 2] - P[a, 1]):
            rows.append(r / np.linalg.norm(r) * w[a])
    return np.array(rows)


def _run(lib, P, uv, w):
    n, V = uv.shape[0], uv.shape[1]
    X = np.zeros((n, 3)); path = np.zeros(n, np.int32); keep = np.ones((n, V), np.uint8)
    lib.hostemu_dlt(n, V, ptr(np.ascontiguousarray(P.reshape(V, 12))), ptr(np.ascontiguousarray(uv)),
                    ptr(np.ascontiguousarray(w)), ptr(keep), 0, ptr(X), ptr(path))
    return X


def test_dlt_matches_lapack_for_shipped_weights():
    lib = hb.load()
    rng = np.random.default_rng(1)
    cams = camera.GetCameraParameters(synth.make_rig("shelf"))
    P = np.stack([c.P for c in cams]).astype(np.float64)
    ages = rng.choice([0, 1, 2, 3], size=(3000, 5))
    ages[:, 0] = 0                                  # a track is only updated when a view matched this frame
    _, uv, w = _systems(rng, P, 3000, ages, 5.0)
    got = _run(lib, P, uv, w)
    for i in range(len(uv)):
        _, _, VT = np.linalg.svd(_rows(P, uv[i], w[i]))
        ref = VT[-1][:3] / VT[-1][3]
        assert np.abs(got[i] - ref).max() < 1e-6     # 1 micrometre; the bar is 0.5 mm


def test_ill_conditioned_system_is_solved_to_working_precision():
    lib = hb.load()
    rng = np.random.default_rng(2)
    cams = camera.GetCameraParameters(synth.make_rig("shelf"))
    P = np.stack([c.P for c in cams]).astype(np.float64)[:2]
    ages = np.tile([3, 0], (400, 1))                 # one fresh view, one 3 frames old, lambda_t = 10
    Xtrue, uv, w = _systems(rng, P, 400, ages, 10.0)
    got = _run(lib, P, uv, w)
    eps = np.finfo(np.float64).eps
    mm_diff = []
    for i in range(len(uv)):
        A = _rows(P, uv[i], w[i])
        _, S, VT = np.linalg.svd(A)
        x_ref = VT[-1]
        x_got = np.r_[got[i], 1.0]; x_got /= np.linalg.norm(x_got)
        if x_got @ x_ref < 0:
            x_got = -x_got
        # (1) it is a smallest singular vector to working precision: ||A x|| is at the rounding floor
        r_got = np.linalg.norm(A.astype(np.longdouble) @ x_got.astype(np.longdouble))
        assert r_got <= S[-1] + 8 * eps * S[0]
        # (2) it differs from LAPACK's vector by no more than the conditioning of the problem allows
        assert np.linalg.norm(x_got - x_ref) <= 8 * eps * S[0] / (S[2] - S[3])
        mm_diff.append(1000 * np.abs(got[i] - x_ref[:3] / x_ref[3]).max())
    # ... which here is millimetres: parity below 0.5 mm is not a meaningful demand in this regime
    assert max(mm_diff) > 0.5
