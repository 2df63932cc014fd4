"""NumPy prototype of the NEXT step of the position-owning bulge chase: deferred rank-2 updates of the
diagonal window D.  Not implemented in CUDA yet; this file pins the algebra down.

In csrc/band_systolic.cu every step reads and rewrites the whole of D in shared memory (D -= v w^H + w v^H): that pass and
y = D v are the shared-memory-bound half of a step.  Here a position keeps the last m pairs (v_i, w_i) instead and works
with D_eff = D - sum_i (v_i w_i^H + w_i v_i^H):

  * y = D_eff v = D v - sum_i [ v_i (w_i^H v) + w_i (v_i^H v) ]               -- 2m dot products and 2m axpys of length b;
  * the column that leaves the window on a slide (and only that one) is brought up to date just in time, like the column
    update of LAPACK's zlatrd; after that the entries of all pending v_i, w_i at the recycled physical slot are zeroed, so
    the row / column that enters there is never touched by updates older than itself;
  * the corner message is D_eff[0, 0] after the step's own pair;
  * after m sweeps the pairs are applied at once, D -= V W^H + W V^H with V, W of b x m: one pass over D instead of m,
    and GEMM-shaped (K = 2m), i.e. work for the FP64 tensor pipe instead of the shared-memory pipe.

Per step, for b = 100 and m = 8: ~6 m b = 4800 element operations on vectors against the 2 b^2 = 20000 element reads and
writes of D they replace; the block update costs 2 b^2 multiply-adds per step amortised, on the DMMA pipe.
The torus addressing makes this simple: v of different sweeps are indexed by global row mod b, so they line up.
"""
import numpy as np

from algo_proto_band import to_band
from algo_proto_systolic import Position


class DeferredPosition(Position):
    def __init__(self, AB, n, b, k, m):
        super().__init__(AB, n, b, k)
        self.m = m
        self.PV = np.zeros((b, 0), complex)       # pending v_i (columns), physical row order
        self.PW = np.zeros((b, 0), complex)
        self.block_updates = 0

    # -- D_eff pieces
    def _deff_matvec(self, v):
        y = self.D @ v
        if self.PV.shape[1]:
            y = y - self.PV @ (self.PW.conj().T @ v) - self.PW @ (self.PV.conj().T @ v)
        return y

    def _deff_column(self, q):
        c = self.D[:, q].copy()
        if self.PV.shape[1]:
            c = c - self.PV @ self.PW[q].conj() - self.PW @ self.PV[q].conj()
        return c

    def _apply_pending(self):
        if self.PV.shape[1]:
            self.D = self.D - self.PV @ self.PW.conj().T - self.PW @ self.PV.conj().T
            phys = np.arange(self.b)
            self.D[phys, phys] = self.D[phys, phys].real
            self.PV = self.PV[:, :0]
            self.PW = self.PW[:, :0]
            self.block_updates += 1

    def step(self, V, TAU):
        # Everything that does not depend on D (reflector, carried block, row message) is the parent's step, run on a
        # scratch D; the D part of the step is restated below on D_eff.
        AB, n, b, k, s, r0, o = self.AB, self.n, self.b, self.k, self.s, self.r0, self.o
        po = (o + b - 1) % b
        D_keep = self.D
        self.D = np.zeros_like(D_keep)             # the parent's D work is discarded
        ab0_keep = AB[0, r0]
        rowmsg = np.zeros(b, complex)
        corner = 0.0
        if r0 + b - 1 < n:
            for j in range(b):
                rowmsg[j] = AB[b - j, r0 - 1 + j]
            corner = AB[0, r0 + b - 1].real
        super().step(V, TAU)                       # advances s, r0, o; writes V, TAU, AB (row message, e); slides Bc
        AB[0, r0] = ab0_keep
        self.D = D_keep
        v = np.zeros(b, complex)
        for p in range(b):
            gr = r0 + (p - o) % b
            if gr < n:
                v[p] = V[gr, s]
        tau = TAU[s, k]
        if k > 0 and min(b, n - r0) <= 1:
            v[:] = 0
            v[o] = 1.0                             # flush step: tau = 0, the stored reflector is zero
        # ... and do the D part on D_eff
        for p in range(b):                         # patch: the entering row / column is current, nothing pending on it
            if p != po:
                val = rowmsg[(p - po) % b]
                self.D[po, p] = val
                self.D[p, po] = np.conj(val)
        self.D[po, po] = corner
        y = tau * self._deff_matvec(v)
        w = y - 0.5 * tau * np.vdot(y, v) * v
        self.PV = np.concatenate([self.PV, v[:, None]], axis=1)
        self.PW = np.concatenate([self.PW, w[:, None]], axis=1)
        col = self._deff_column(o)                 # the column that leaves, brought up to date just in time
        AB[0, r0] = col[o].real                    # corner message
        # slide: the parent slid with its scratch D; redo with the true column
        if k > 0:
            self.Bc[:, o] = col
            self.Bc[o, :] = 0
        else:
            self.xcol = col.copy()
            self.xcol[o] = 0
        self.PV[o, :] = 0                          # the physical slot is recycled: older updates must not reach its
        self.PW[o, :] = 0                          # next occupant
        if self.PV.shape[1] == self.m:
            self._apply_pending()


def chase_systolic_deferred(A, b, m):
    n = A.shape[0]
    AB = to_band(A, b)
    V = np.zeros((n, n), complex)
    TAU = np.zeros((n, (n + b - 1) // b + 1), complex)
    KP = (n - 2) // b + 1
    pos = [DeferredPosition(AB, n, b, k, m) for k in range(KP)]
    for s in range(n - 1):
        for k in range(KP):
            if pos[k].active():
                pos[k].step(V, TAU)
    return AB[0, :].real.copy(), AB[1, :n - 1].real.copy(), V, TAU, sum(p.block_updates for p in pos)
