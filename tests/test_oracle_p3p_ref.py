"""oracle/pose_oracle.cpp's P3P restatement against the UNMODIFIED reference p3p.cpp (oracle/_ref/libref_p3p.so, built by
oracle/Makefile from /root/reference against the small Eigen stand-in in oracle/eigen_shim/): bit-for-bit on random,
degenerate and inconsistent problems.  Skipped when the library was not built (no /root/reference at build time)."""
import ctypes as C
import os

import numpy as np
import pytest

from rpg_monocular_pose_estimator_b200 import synth
from oracle import pose_oracle

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_p3p.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libref_p3p.so not built")


def _ref():
    L = C.CDLL(REF)
    dp = C.POINTER(C.c_double)
    L.ref_p3p.argtypes = [dp, dp, dp]; L.ref_p3p.restype = C.c_int
    L.ref_solve_quartic.argtypes = [dp, dp]; L.ref_solve_quartic.restype = C.c_int
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_p3p_bit_identical_to_reference_source():
    L = _ref()
    rng = np.random.default_rng(123)
    n_checked = 0
    for i in range(3000):
        pts = rng.uniform(-0.2, 0.2, size=(3, 3))
        if i % 40 == 0:
            pts[2] = pts[0] + 1.5 * (pts[1] - pts[0])                     # colinear
        Rm = synth.rodrigues(rng.normal(size=3) * 0.9)
        t = np.array([rng.uniform(-.3, .3), rng.uniform(-.3, .3), rng.uniform(.3, 1.5)])
        cam = (Rm @ pts.T).T + t
        f = cam / np.linalg.norm(cam, axis=1, keepdims=True)
        if i % 3 == 0:
            f = f + rng.normal(size=f.shape) * rng.choice([1e-3, 0.05, 0.5]); f /= np.linalg.norm(f, axis=1, keepdims=True)
        fc = np.ascontiguousarray(f); pc = np.ascontiguousarray(pts)       # column k of the 3x3 = row k here, stored contiguously
        sol_ref = np.zeros(48)
        rc_ref = L.ref_p3p(_dp(fc), _dp(pc), _dp(sol_ref))
        rc, sol = pose_oracle.p3p(f.T, pts.T)
        assert rc == rc_ref
        a, b = sol.reshape(-1), sol_ref
        assert np.array_equal(np.isnan(a), np.isnan(b))
        m = ~np.isnan(a)
        assert np.array_equal(a[m], b[m]), (i, np.abs(a[m] - b[m]).max())
        n_checked += 1
    assert n_checked == 3000


def test_quartic_bit_identical_to_reference_source():
    L = _ref()
    rng = np.random.default_rng(7)
    for i in range(2000):
        fac = rng.normal(size=5) * 10.0 ** rng.integers(-3, 4, size=5)
        if i % 10 == 0:
            fac[1] = 0.0
        if i % 17 == 0:
            r = rng.normal(size=4); fac = np.poly(r)                       # four real roots
        a = pose_oracle.solve_quartic(fac)
        b = np.zeros(4); L.ref_solve_quartic(_dp(np.ascontiguousarray(fac, np.float64)), _dp(b))
        assert np.array_equal(np.isnan(a), np.isnan(b))
        assert np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
