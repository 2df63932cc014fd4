"""NumPy prototype of the band route of the eigensolver (csrc/band.cu): the formulas and the
application order the CUDA kernels implement, checked against LAPACK in test_algo_proto.py.

  fold ordering        sites of each ring as 0, L-1, 1, L-2, ...; particle / hole interleaved
                       -> the BdG matrix is a band matrix of half-bandwidth 4 min(Lx, Ly) + 4
  chase_band           Householder bulge chasing on lower band storage AB[d, j] = A[j + d, j]
                       (LD = 2b): sweep s, step k, reflector on rows s+1+kb .. s+(k+1)b; the block
                       pushed below the band is carried to the next step with its right-application
                       deferred, exactly as the kernel keeps it in shared memory
  backtransform_blocked  U = Q2 Z with the reflectors of g consecutive sweeps at one step merged
                       into one staircase block reflector; groups last to first, steps ascending
"""
import numpy as np


def fold_positions(L):
    order, lo, hi = [], 0, L - 1
    while lo <= hi:
        order.append(lo)
        if hi != lo:
            order.append(hi)
        lo += 1
        hi -= 1
    pos = np.empty(L, int)
    pos[order] = np.arange(L)
    return pos


def band_positions(Lx, Ly):
    """pos[r]: band index of row r (r < N particle of site r, r >= N hole), short ring fastest."""
    px, py = fold_positions(Lx), fold_positions(Ly)
    N = Lx * Ly
    pos = np.empty(2 * N, int)
    for y in range(Ly):
        for x in range(Lx):
            i = y * Lx + x
            s = py[y] * Lx + px[x] if Lx <= Ly else px[x] * Ly + py[y]
            pos[i], pos[i + N] = 2 * s, 2 * s + 1
    return pos


def larfg(x):
    """LAPACK zlarfg: H^H x = beta e1 with H = I - tau v v^H, v[0] = 1, beta real."""
    alpha = x[0]
    xn = np.linalg.norm(x[1:]) if len(x) > 1 else 0.0
    if xn == 0.0 and alpha.imag == 0.0:
        v = np.zeros_like(x)
        v[0] = 1
        return v, 0.0, alpha.real
    beta = -np.copysign(np.sqrt(abs(alpha) ** 2 + xn ** 2), alpha.real)
    tau = complex((beta - alpha.real) / beta, -alpha.imag / beta)
    v = x / (alpha - beta)
    v[0] = 1
    return v, tau, beta


def to_band(A, b):
    n, LD = A.shape[0], 2 * b
    AB = np.zeros((LD, n), complex)
    for j in range(n):
        m = min(LD, n - j)
        AB[:m, j] = A[j:j + m, j]
    return AB


def _get(AB, r0, nr, c0, nc):
    LD = AB.shape[0]
    B = np.zeros((nr, nc), complex)
    for jj in range(nc):
        for ii in range(nr):
            d = r0 + ii - (c0 + jj)
            if 0 <= d < LD:
                B[ii, jj] = AB[d, c0 + jj]
    return B


def _put(AB, B, r0, c0):
    LD = AB.shape[0]
    for jj in range(B.shape[1]):
        for ii in range(B.shape[0]):
            d = r0 + ii - (c0 + jj)
            if 0 <= d < LD:
                AB[d, c0 + jj] = B[ii, jj]


def chase_band(A, b, dense_tail=False):
    """Hermitian A of half-bandwidth b -> (d, e, V, TAU): real tridiagonal (d, e); V[:, s] holds the
    reflectors of sweep s at their row positions, TAU[s, k] their tau.  dense_tail: the last b sweeps
    (one step each) as a dense tridiagonalisation of the trailing (b+1) x (b+1) block, the way
    chase_tail_kernel does them in shared memory."""
    n = A.shape[0]
    AB = to_band(A, b)
    V = np.zeros((n, n), complex)
    TAU = np.zeros((n, (n + b - 1) // b + 1), complex)
    nsweep = n - 1 - b if dense_tail and n - 1 - b >= 1 else n - 1
    for s in range(nsweep):
        k, r0 = 0, s + 1
        Bc = us = vp = None
        taup = 0.0
        while True:
            ln = min(b, n - r0)
            if k > 0 and ln <= 1:          # nothing to annihilate: flush the carried block
                _put(AB, Bc - taup * np.outer(us, vp.conj()), r0, r0 - b)
                break
            x = AB[1:1 + ln, s].copy() if k == 0 else Bc[:, 0] - taup * us      # vp[0] = 1
            v, tau, beta = larfg(x)
            V[r0:r0 + ln, s] = v
            TAU[s, k] = tau
            if k == 0:
                AB[1, s] = beta
                AB[2:1 + ln, s] = 0
            else:                          # pending right-application + left-application of H^H
                tu = taup * us
                z = v.conj() @ Bc - (v.conj() @ tu) * vp.conj()
                new = Bc - np.outer(tu, vp.conj()) - np.conj(tau) * np.outer(v, z)
                new[:, 0] = 0
                new[0, 0] = beta
                _put(AB, new, r0, r0 - b)
            D = _get(AB, r0, ln, r0, ln)   # diagonal block, two-sided (zhetd2 formulas)
            Df = np.tril(D) + np.tril(D, -1).conj().T
            xw = tau * (Df @ v)
            w = xw - 0.5 * tau * np.vdot(xw, v) * v
            _put(AB, np.tril(Df - np.outer(v, w.conj()) - np.outer(w, v.conj())), r0, r0)
            r1 = r0 + ln
            if r1 >= n:
                break
            Bc = _get(AB, r1, min(b, n - r1), r0, ln)
            us, vp, taup = Bc @ v, v, tau
            r0, k = r1, k + 1
    if nsweep < n - 1:
        s0, m0 = n - 1 - b, b + 1
        F = _get(AB, s0, m0, s0, m0)       # trailing block: full storage from the lower band
        F = np.tril(F) + np.tril(F, -1).conj().T
        for j in range(b):
            v, tau, beta = larfg(F[j + 1:, j].copy())
            V[s0 + j + 1:, s0 + j] = v
            TAU[s0 + j, 0] = tau
            A22 = F[j + 1:, j + 1:]
            y = tau * (A22 @ v)
            w = y - 0.5 * tau * np.vdot(y, v) * v
            F[j + 1:, j + 1:] = A22 - np.outer(v, w.conj()) - np.outer(w, v.conj())
            F[j + 1:, j] = 0
            F[j + 1, j] = beta
        for c in range(m0):
            AB[0, s0 + c] = F[c, c].real
            if c + 1 < m0:
                AB[1, s0 + c] = F[c + 1, c].real
    return AB[0, :].real.copy(), AB[1, :n - 1].real.copy(), V, TAU


def backtransform_blocked(Z, V, TAU, b, g):
    """U = Q2 Z.  Block (s0, k): columns s0..s0+g-1 of V, rows s0+1+kb..; column c non-zero on rows [c, c+b)."""
    n = Z.shape[0]
    Z = Z.astype(complex).copy()
    ngrp = (n - 1 + g - 1) // g
    for G in range(ngrp - 1, -1, -1):
        s0 = G * g
        gg = min(g, n - 1 - s0)
        k = 0
        while True:
            rlo = s0 + 1 + k * b
            if n - rlo < 1 or (k > 0 and n - rlo < 2):
                break
            rows = min(n - rlo, b + gg - 1)
            Vb = np.zeros((rows, gg), complex)
            for c in range(gg):
                lo, hi = c, min(c + b, rows)
                Vb[lo:hi, c] = V[rlo + lo:rlo + hi, s0 + c]
            T = np.zeros((gg, gg), complex)
            for i in range(gg):
                T[i, i] = TAU[s0 + i, k]
                if i > 0:
                    T[:i, i] = -TAU[s0 + i, k] * (T[:i, :i] @ (Vb[:, :i].conj().T @ Vb[:, i]))
            Z[rlo:rlo + rows] -= Vb @ (T @ (Vb.conj().T @ Z[rlo:rlo + rows]))
            k += 1
    return Z


def backtransform_wavefronts(Z, V, TAU, b, g):
    """Same product with the blocks applied wavefront by wavefront, t = (ngrp - 1 - G) + k (the launch order of
    band_apply_kernel): blocks of one wavefront touch disjoint rows, every dependency has a smaller t."""
    n = Z.shape[0]
    Z = Z.astype(complex).copy()
    ngrp = (n - 1 + g - 1) // g
    blocks = []
    for G in range(ngrp):
        s0 = G * g
        k = 0
        while True:
            rlo = s0 + 1 + k * b
            if n - rlo < 1 or (k > 0 and n - rlo < 2):
                break
            blocks.append(((ngrp - 1 - G) + k, G, k))
            k += 1
    for t in sorted({blk[0] for blk in blocks}):
        wave = [blk for blk in blocks if blk[0] == t]
        touched = []
        for _, G, k in wave:
            s0 = G * g
            gg = min(g, n - 1 - s0)
            rlo = s0 + 1 + k * b
            rows = min(n - rlo, b + gg - 1)
            assert all(rlo + rows <= lo or hi <= rlo for lo, hi in touched), "blocks of a wavefront overlap"
            touched.append((rlo, rlo + rows))
            Vb = np.zeros((rows, gg), complex)
            for c in range(gg):
                lo, hi = c, min(c + b, rows)
                Vb[lo:hi, c] = V[rlo + lo:rlo + hi, s0 + c]
            T = np.zeros((gg, gg), complex)
            for i in range(gg):
                T[i, i] = TAU[s0 + i, k]
                if i > 0:
                    T[:i, i] = -TAU[s0 + i, k] * (T[:i, :i] @ (Vb[:, :i].conj().T @ Vb[:, i]))
            Z[rlo:rlo + rows] -= (Vb @ T) @ (Vb.conj().T @ Z[rlo:rlo + rows])
    return Z

