"""NumPy prototype of the blocked Hermitian tridiagonalisation / back-transformation that
csrc/hetrd.cu and csrc/backtransform.cu implement (TESTS ONLY: it documents the exact
formulas -- conjugations, signs, index ranges -- the CUDA kernels follow, and is checked
against LAPACK in tests/test_algo_proto.py).  0-based, lower storage, reflectors
H_j = I - tau_j v_j v_j^H with v_j[j+1] = 1, Q = H_0 H_1 ... H_{n-2}, A = Q T Q^H."""
import numpy as np


def larfg(alpha, x):
    """Return (beta, tau, scale): H^H [alpha; x] = [beta; 0], v = [1; x*scale]."""
    xn2 = float(np.sum(np.abs(x) ** 2))
    if xn2 == 0.0 and alpha.imag == 0.0:
        return alpha.real, 0.0 + 0.0j, 0.0 + 0.0j
    beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn2), alpha.real)
    tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
    scale = 1.0 / (alpha - beta)
    return beta, tau, scale


def hetrd_blocked(A_in, nb=32):
    n = A_in.shape[0]
    A = A_in.copy()                        # full storage, both triangles kept valid
    V = np.zeros((n, n), complex)          # column j = reflector j (rows > j), explicit unit
    W = np.zeros((n, n), complex)
    d = np.zeros(n)
    e = np.zeros(max(n - 1, 0))
    tau = np.zeros(max(n - 1, 0), complex)
    P2 = np.zeros((n, nb), complex)        # P2[j, k] = V[:, j0+k]^H v_j (k < i): feeds larft
    j0 = 0
    while j0 < n - 1:
        pn = min(nb, n - 1 - j0)
        for i in range(pn):
            j = j0 + i
            Vp, Wp = V[:, j0:j0 + i], W[:, j0:j0 + i]
            # --- make_reflector(j)
            a = A[j:, j] - Vp[j:] @ Wp[j].conj() - Wp[j:] @ Vp[j].conj()
            d[j] = a[0].real
            beta, t, scale = larfg(a[1], a[2:])
            e[j] = beta
            tau[j] = t
            v = np.zeros(n, complex)
            v[j + 1] = 1.0
            v[j + 2:] = a[2:] * scale
            V[:, j] = v
            p1 = Wp[j + 1:].conj().T @ v[j + 1:]
            p2 = Vp[j + 1:].conj().T @ v[j + 1:]
            P2[j, :i] = p2
            # --- hemv(j)
            y = A[j + 1:, j + 1:] @ v[j + 1:]
            # --- finish_w(j)
            w = y - Vp[j + 1:] @ p1 - Wp[j + 1:] @ p2
            w = t * w
            dot = np.vdot(w, v[j + 1:])
            w = w - 0.5 * t * dot * v[j + 1:]
            W[j + 1:, j] = w
        j1 = j0 + pn
        Vp, Wp = V[j1:, j0:j1], W[j1:, j0:j1]
        A[j1:, j1:] -= Vp @ Wp.conj().T + Wp @ Vp.conj().T
        j0 = j1
    d[n - 1] = A[n - 1, n - 1].real
    return d, e, tau, V, P2


def larft_blocks(V, tau, P2, nb=32):
    """T factors (forward, columnwise) per block from the saved V^H v products."""
    n = V.shape[0]
    Ts = []
    for j0 in range(0, n - 1, nb):
        pn = min(nb, n - 1 - j0)
        T = np.zeros((nb, nb), complex)
        for i in range(pn):
            T[i, i] = tau[j0 + i]
            if i:
                T[:i, i] = -tau[j0 + i] * (T[:i, :i] @ P2[j0 + i, :i])
        Ts.append(T)
    return Ts


def backtransform(V, Ts, Z, nb=32):
    n = V.shape[0]
    U = Z.astype(complex).copy()
    nblk = len(Ts)
    for k in range(nblk - 1, -1, -1):
        j0 = k * nb
        pn = min(nb, n - 1 - j0)
        Vk = V[j0 + 1:, j0:j0 + pn]
        W1 = Vk.conj().T @ U[j0 + 1:]
        W2 = Ts[k][:pn, :pn] @ W1
        U[j0 + 1:] -= Vk @ W2
    return U
