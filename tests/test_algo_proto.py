"""The NumPy prototype of the blocked tridiagonalisation / back-transform formulas used by the
CUDA eigensolver reproduces LAPACK eigenpairs (formula check, CPU only)."""
import numpy as np
import pytest
import scipy.linalg as sl

from algo_proto import backtransform, hetrd_blocked, larft_blocks


@pytest.mark.parametrize("n,nb", [(2, 32), (3, 4), (33, 32), (70, 8), (130, 32)])
def test_blocked_hetrd_backtransform(n, nb):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = A + A.conj().T
    d, e, tau, V, P2 = hetrd_blocked(A, nb)
    w, Z = sl.eigh_tridiagonal(d, e)
    wr = np.linalg.eigvalsh(A)
    assert np.max(np.abs(w - wr)) < 1e-12 * max(1, np.max(np.abs(wr)))
    U = backtransform(V, larft_blocks(V, tau, P2, nb), Z, nb)
    assert np.max(np.abs(U.conj().T @ U - np.eye(n))) < 1e-12
    assert np.max(np.abs(A @ U - U * w)) < 1e-11
