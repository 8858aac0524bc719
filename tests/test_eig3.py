"""CPU: the product's dgeev-faithful 3x3 principal axis (csrc/octa_eig3.h, host build of the same
host/device code) against the LAPACK the reference calls (np.linalg.eig, greenhouse.py:229-233)."""
import ctypes

import numpy as np
import scipy.linalg as sl
import scipy.linalg.lapack as LP

from octa_autosegmentation_b200 import _lib


def _cov(rng, m):
    pts = rng.uniform(0, 1, (m, 3)) * np.array([0.1, 0.1, 0.0131]) + rng.uniform(0, 1, 3)
    return np.cov((pts - pts.mean(axis=0)).T)


def test_principal_axis_sign_and_value_match_numpy():
    L = _lib.lib()
    rng = np.random.RandomState(7)
    n_complex = 0
    for trial in range(6000):
        m = 2 if trial % 3 == 0 else int(rng.randint(2, 40))
        cov = np.ascontiguousarray(_cov(rng, m))
        w, v = np.linalg.eig(cov)
        n_complex += np.iscomplexobj(w)
        ref = np.real(v[:, int(np.argmax(w))])
        got = np.zeros(3)
        assert L.octa_test_principal_axis(cov.ctypes.data, got.ctypes.data) == 0
        assert np.abs(got - ref).max() < 1e-13, (m, got, ref)
    assert n_complex > 0      # the complex-typed corner (two ~zero eigenvalues) is exercised


def test_schur_factors_bit_identical_to_lapack():
    """The sign is decided by the parity of QR sweeps, which hinges on rounding-level noise in the
    deflation test -- so Hessenberg form, Q, and the Schur pair (T, Z) must match LAPACK bit for bit."""
    L = _lib.lib()
    rng = np.random.RandomState(11)
    for trial in range(1500):
        cov = np.ascontiguousarray(_cov(rng, int(rng.randint(3, 12))))
        w3, v9, dbg = np.zeros(3), np.zeros(9), np.zeros(48)
        st = L.octa_test_eig3_debug(cov.ctypes.data, w3.ctypes.data, v9.ctypes.data, dbg.ctypes.data)
        ht, tau, _ = LP.dgehrd(cov, lo=0, hi=2)
        q, _ = LP.dorghr(ht, tau, lo=0, hi=2)
        assert np.array_equal(np.triu(ht, -1), dbg[:9].reshape(3, 3))
        assert np.array_equal(q, dbg[9:18].reshape(3, 3))
        if st == 0:
            T, Z = sl.schur(cov)
            assert np.array_equal(T, dbg[18:27].reshape(3, 3)) and np.array_equal(Z, dbg[27:36].reshape(3, 3))
