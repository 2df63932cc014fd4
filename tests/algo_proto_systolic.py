"""NumPy prototype of the position-owning ("systolic") bulge chase of csrc/band_systolic.cu.

Same reflectors, tau and tridiagonal matrix as algo_proto_band.chase_band, but organised the way the kernel is:
position k of a chain owns step k of EVERY sweep.  Its two b x b windows -- the carried block Bc (rows r0..r0+b-1,
columns r0-b..r0-1) and the diagonal block D (rows / columns r0..r0+b-1), r0 = s + 1 + k b -- slide down the
diagonal by one per sweep and stay on chip: element (global row g, global column c) lives at the physical slot
(g mod b, c mod b) of the window, so a slide overwrites one physical row and one physical column and nothing
else moves.  Rows / columns beyond the matrix are kept at zero, so no length masks are needed in the block
operations.  Per sweep a position exchanges O(b) numbers with its neighbours, all through their true places in
the band storage AB (no separate mailboxes):
  k-1 -> k   the reflector v(s, k-1) and its tau (the V / TAU outputs themselves)
  k+1 -> k   row 0 of Bc(s-1, k+1) after its update (-> new last row of D(s, k), corner of Bc(s, k)) and the
             corner D(s-1, k+1)[0, 0]
The column that leaves D on a slide becomes the new last column of Bc (position 0: the next column to annihilate).
"""
import numpy as np

from algo_proto_band import larfg, to_band


class Position:
    def __init__(self, AB, n, b, k, s0=0):
        self.AB, self.n, self.b, self.k = AB, n, b, k
        self.s = s0
        self.r0 = s0 + 1 + k * b
        self.o = self.r0 % b
        b_, r0, o = b, self.r0, self.o
        self.Bc = np.zeros((b_, b_), complex)
        self.D = np.zeros((b_, b_), complex)
        self.xcol = np.zeros(b_, complex)
        po = (o + b_ - 1) % b_
        for p in range(b_):                       # physical row
            gr = r0 + (p - o) % b_
            for q in range(b_):                   # physical column
                if gr >= n:
                    continue
                if k > 0 and p != po:
                    gc = r0 - b_ + (q - o) % b_
                    self.Bc[p, q] = AB[gr - gc, gc]
                gc = r0 + (q - o) % b_
                if gc < n and p != po and q != po:
                    self.D[p, q] = AB[gr - gc, gc] if gr >= gc else np.conj(AB[gc - gr, gr])
                    if gr == gc:
                        self.D[p, q] = self.D[p, q].real
            if k == 0 and gr < n and p != po:
                self.xcol[p] = AB[gr - s0, s0]

    def active(self):
        return self.r0 <= self.n - 1

    def lg(self, p):
        return (p - self.o) % self.b

    def step(self, V, TAU):
        AB, n, b, k, s, r0, o = self.AB, self.n, self.b, self.k, self.s, self.r0, self.o
        po = (o + b - 1) % b
        ln = min(b, n - r0)
        phys = np.arange(b)
        # ---- helper: messages of this step (zeros where the matrix ends)
        rowmsg = np.zeros(b, complex)
        corner = 0.0
        if r0 + b - 1 < n:
            for j in range(b):
                rowmsg[j] = AB[b - j, r0 - 1 + j]          # element (r0+b-1, r0-1+j)
            corner = AB[0, r0 + b - 1].real
        if k > 0:
            vp = np.zeros(b, complex)
            for p in range(b):
                vp[p] = V[r0 - b + (p - o) % b, s]           # columns always exist
            taup = TAU[s, k - 1]
            # ---- a. u = Bc vp (the new last row still zero), then the corner
            u = self.Bc @ vp
            u[po] += rowmsg[0] * vp[po]
            self.Bc[po, po] = rowmsg[0]
            tu = taup * u
            x = self.Bc[:, o] - tu
        else:
            x = self.xcol.copy()
            x[po] = rowmsg[0]
        # ---- b. reflector (logical order for the norm: physical row o is logical 0)
        if k > 0 and ln <= 1:
            tau, beta, v = 0.0, x[o], np.zeros(b, complex)
            v[o] = 1.0
            vout = np.zeros(b, complex)
        else:
            xl = np.array([x[(o + l) % b] for l in range(ln)])
            vl, tau, beta = larfg(xl)
            v = np.zeros(b, complex)
            for l in range(ln):
                v[(o + l) % b] = vl[l]
            vout = v
        for p in range(b):
            gr = r0 + (p - o) % b
            if gr < n:
                V[gr, s] = vout[p]
        TAU[s, k] = tau
        if k == 0:
            AB[1, s] = beta
        else:
            # ---- c. carried block
            z = v.conj() @ self.Bc - (v.conj() @ tu) * vp.conj()
            wc = np.conj(tau) * z
            self.Bc = self.Bc - np.outer(tu, vp.conj()) - np.outer(v, wc)
            self.Bc[:, o] = 0
            self.Bc[o, o] = beta
            for q in range(b):                            # row message: global row r0
                gc = r0 - b + (q - o) % b
                AB[r0 - gc, gc] = self.Bc[o, q]
        # ---- d. diagonal block: patch the new row / column, two-sided update
        for p in range(b):
            if p != po:
                val = rowmsg[(p - po) % b]
                self.D[po, p] = val
                self.D[p, po] = np.conj(val)
        self.D[po, po] = corner
        y = tau * (self.D @ v)
        w = y - 0.5 * tau * np.vdot(y, v) * v
        AB[0, r0] = (self.D[o, o] - v[o] * np.conj(w[o]) - w[o] * np.conj(v[o])).real   # corner message
        self.D = self.D - np.outer(v, w.conj()) - np.outer(w, v.conj())
        self.D[phys, phys] = self.D[phys, phys].real
        # ---- slide
        if k > 0:
            self.Bc[:, o] = self.D[:, o]
            self.Bc[o, :] = 0
        else:
            self.xcol = self.D[:, o].copy()
            self.xcol[o] = 0
        self.s += 1
        self.r0 += 1
        self.o = (o + 1) % b


def chase_systolic(A, b):
    n = A.shape[0]
    AB = to_band(A, b)
    V = np.zeros((n, n), complex)
    TAU = np.zeros((n, (n + b - 1) // b + 1), complex)
    KP = (n - 2) // b + 1
    pos = [Position(AB, n, b, k) for k in range(KP)]
    for s in range(n - 1):
        for k in range(KP):
            if pos[k].active():
                assert pos[k].s == s
                pos[k].step(V, TAU)
    return AB[0, :].real.copy(), AB[1, :n - 1].real.copy(), V, TAU
