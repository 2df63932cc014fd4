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


def test_band_route_prototype_matches_lapack():
    """tests/algo_proto_band.py (the formulas of csrc/band.cu): fold ordering -> band matrix of
    half-bandwidth 4L+4; bulge chase -> tridiagonal with the same spectrum; blocked staircase
    back-transformation -> eigenvectors."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import dwhmc_oracle as orc
    import scipy.linalg as sl
    import algo_proto_band as apb
    for L, g in ((6, 5), (8, 29)):
        p = orc.ModelParameters(L, L, 1.0, -0.35, -1.08, 1.0, 0.05, 20.0, 0.8, 1.0)
        _, st, c = orc.make_chain(p, 3)
        H = orc.full_hermitian(c)
        n = 2 * L * L
        pos = apb.band_positions(L, L)
        Hp = np.zeros_like(H)
        Hp[np.ix_(pos, pos)] = H
        i, j = np.nonzero(Hp)
        b = int(np.max(np.abs(i - j)))
        assert b == 4 * L + 4
        d, e, V, TAU = apb.chase_band(Hp, b)
        w, Z = sl.eigh_tridiagonal(d, e)
        assert np.max(np.abs(w - c.E_n)) <= 1e-12 * np.max(np.abs(c.E_n))
        U = apb.backtransform_blocked(Z, V, TAU, b, g)
        assert np.max(np.abs(Hp @ U - U * w)) <= 1e-12 * np.max(np.abs(w))
        assert np.max(np.abs(U.conj().T @ U - np.eye(n))) <= 1e-12
        Uw = apb.backtransform_wavefronts(Z, V, TAU, b, g)          # the launch order of the fused kernel
        assert np.max(np.abs(Uw - U)) <= 1e-13
        # the last b sweeps as a dense tridiagonalisation of the trailing block (chase_tail_kernel): same tridiagonal
        # matrix, reflectors and tau as the bulge chase produces for those single-step sweeps
        d2, e2, V2, TAU2 = apb.chase_band(Hp, b, dense_tail=True)
        scale = np.max(np.abs(w))
        assert np.max(np.abs(d2 - d)) <= 1e-12 * scale and np.max(np.abs(e2 - e)) <= 1e-12 * scale
        assert np.max(np.abs(V2 - V)) <= 1e-11 and np.max(np.abs(TAU2 - TAU)) <= 1e-11


def test_systolic_chase_prototype_matches_sweep_owning_prototype():
    """tests/algo_proto_systolic.py (the organisation of csrc/band_systolic.cu: a position owns step k of every
    sweep; windows slide on a torus, messages travel through the band storage) produces the same tridiagonal
    matrix, reflectors and tau as the sweep-owning prototype, on matrices whose order is / is not a multiple of the
    half-bandwidth, with one position only, and with a band as wide as the matrix allows."""
    import algo_proto_band as apb
    import algo_proto_systolic as aps
    rng = np.random.default_rng(1)
    for n, b in ((23, 4), (40, 7), (37, 5), (30, 28), (64, 12), (50, 3), (72, 28)):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        A = A + A.conj().T
        i, j = np.indices((n, n))
        A[np.abs(i - j) > b] = 0
        d0, e0, V0, T0 = apb.chase_band(A, b)
        d1, e1, V1, T1 = aps.chase_systolic(A, b)
        scale = np.max(np.abs(A))
        assert np.max(np.abs(d0 - d1)) <= 1e-12 * scale and np.max(np.abs(e0 - e1)) <= 1e-12 * scale
        assert np.max(np.abs(V0 - V1)) <= 1e-11 and np.max(np.abs(T0 - T1)) <= 1e-11


def test_deferred_update_prototype_matches_systolic_prototype():
    """tests/algo_proto_deferred.py (the next step for the chase kernel, DESIGN.md section 7: the rank-2 updates of the
    diagonal window are kept as m pending pairs, the leaving column is refreshed just in time, the pairs are applied as
    one GEMM-shaped block update every m sweeps) gives the same tridiagonal matrix, reflectors and tau as the
    step-by-step prototype of the kernel as it is."""
    import algo_proto_systolic as aps
    import algo_proto_deferred as apd
    rng = np.random.default_rng(2)
    for n, b, m in ((23, 4, 3), (40, 7, 4), (37, 5, 8), (64, 12, 8), (30, 28, 5), (50, 3, 2), (45, 9, 1)):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        A = A + A.conj().T
        i, j = np.indices((n, n))
        A[np.abs(i - j) > b] = 0
        d1, e1, V1, T1 = aps.chase_systolic(A, b)
        d2, e2, V2, T2, nblock = apd.chase_systolic_deferred(A, b, m)
        scale = np.max(np.abs(A))
        assert np.max(np.abs(d1 - d2)) <= 1e-12 * scale and np.max(np.abs(e1 - e2)) <= 1e-12 * scale
        assert np.max(np.abs(V1 - V2)) <= 1e-11 and np.max(np.abs(T1 - T2)) <= 1e-11
        assert nblock > 0
