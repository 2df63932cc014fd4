"""CPU check of csrc/stedc_core.h (secular solver, deflation, QL leaf) through the
tests-only host harness tests/hostcheck/stedc_host.cpp, against LAPACK."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sl

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    d = os.path.join(HERE, "hostcheck")
    subprocess.check_call(["make", "-s", "-C", d])
    return ctypes.CDLL(os.path.join(d, "libhostcheck.so"))


def stedc(lib, d, e, leaf=36):
    n = len(d)
    w = np.zeros(n); Z = np.zeros((n, n)); st = (ctypes.c_int * 4)()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    d = np.ascontiguousarray(d, dtype=float); e = np.ascontiguousarray(e, dtype=float)
    rc = lib.host_stedc(n, P(d), P(e), P(w), P(Z), leaf, st)
    return rc, w, Z.T.copy(), list(st)


def cases():
    rng = np.random.default_rng(0)
    out = []
    for n in (1, 2, 3, 37, 73, 128, 200, 512):
        out.append((f"random{n}", rng.standard_normal(n), rng.standard_normal(max(n - 1, 0))))
    n = 257
    out.append(("wilkinson", np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1)))
    out.append(("toeplitz", 2 * np.ones(300), -np.ones(299)))
    out.append(("zeros", np.zeros(100), np.zeros(99)))
    out.append(("tiny_e", rng.standard_normal(200), 1e-14 * rng.standard_normal(199)))
    d = rng.standard_normal(300); e = rng.standard_normal(299); e[::7] = 0
    out.append(("split", d, e))
    out.append(("clustered", 1 + 1e-10 * rng.standard_normal(300), 1e-10 * rng.standard_normal(299)))
    out.append(("graded", 10.0 ** (-np.arange(200) / 20), 10.0 ** (-np.arange(199) / 20)))
    W = np.abs(np.arange(21) - 10).astype(float); d = np.tile(W, 10); e = np.ones(len(d) - 1); e[20::21] = 1e-9
    out.append(("glued_wilkinson", d, e))
    return out


@pytest.mark.parametrize("name,d,e", cases(), ids=[c[0] for c in cases()])
def test_host_dc_matches_lapack(lib, name, d, e):
    rc, w, Z, st = stedc(lib, d, e)
    assert rc == 0 and st[1] == 0
    n = len(d)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    wr = sl.eigh_tridiagonal(d, e, eigvals_only=True) if n > 1 else d
    nrm = max(np.max(np.abs(wr)), 1e-300)
    assert np.max(np.abs(w - wr)) <= 5e-14 * nrm
    assert np.max(np.abs(Z.T @ Z - np.eye(n))) <= 5e-14
    assert np.max(np.abs(T @ Z - Z * w)) <= 5e-14 * nrm
    assert np.all(np.diff(w) >= 0)
