// stedc_core.h -- scalar numerics of the tridiagonal divide-and-conquer eigensolver,
// written once for host and device (the CUDA kernels in stedc.cu call these with a
// warp-parallel reduction policy; tests/hostcheck compiles them with g++ and a serial
// policy so the delicate logic is unit-tested on the CPU box).
//
// Algorithm: Cuppen's divide and conquer with Gu/Eisenstat's stable eigenvector
// formula (the method behind LAPACK dstedc/dlaed0-4; restated from the published
// algorithm, not from LAPACK source).  It replaces the tridiagonal stage of the
// reference's eigen!(Hermitian(U,:U)) call (src/Hamiltonian.jl:106).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define DW_HD __host__ __device__ __forceinline__
#else
#define DW_HD inline
#endif

namespace dwcore {

// relative machine precision as LAPACK defines it (2^-53)
#define DW_EPS 1.1102230246251565e-16

// ---------------------------------------------------------------------------
// Serial reduction policy (host, or one device thread).  The device warp policy in
// stedc.cu has the same interface: begin/stride partition an index range and sum4
// returns the totals to every participant.
// ---------------------------------------------------------------------------
struct SerialPar {
  DW_HD int begin() const { return 0; }
  DW_HD int stride() const { return 1; }
  DW_HD void sum4(double&, double&, double&, double&) const {}
  DW_HD void sum1(double&) const {}
};

// ---------------------------------------------------------------------------
// Implicit-shift QL iteration on a small symmetric tridiagonal matrix (leaf solver).
// d[0..n), e[0..n-1) (e[n-1] is scratch).  rot(i, c, s) must apply the plane rotation
//   z[:,i+1] <- s*z[:,i] + c*z[:,i+1] ;  z[:,i] <- c*z[:,i] - s*z[:,i+1]
// to the accumulated eigenvector matrix.  Returns 0, or l+1 if eigenvalue l failed to
// converge in 60 sweeps.
// ---------------------------------------------------------------------------
template <class Rot>
DW_HD int tql_implicit(int n, double* d, double* e, Rot& rot) {
  if (n <= 1) return 0;
  e[n - 1] = 0.0;
  for (int l = 0; l < n; ++l) {
    int iter = 0;
    int m;
    do {
      for (m = l; m < n - 1; ++m) {
        double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= DW_EPS * dd) break;
      }
      if (m != l) {
        if (iter++ == 60) return l + 1;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          double b = c * e[i];
          // sqrt(f^2 + g^2) directly where that cannot over- or underflow (hypot's scaling costs about as much as the
          // rest of the rotation, and this recurrence is one long dependency chain per leaf)
          const double r2 = f * f + g * g;
          r = (r2 > 1e-280 && r2 < 1e280) ? sqrt(r2) : hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          const double rinv = 1.0 / r;
          s = f * rinv;
          c = g * rinv;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          rot(i, c, s);
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return 0;
}

// ---------------------------------------------------------------------------
// Deflation scan of one merge (sequential; cf. the published description of dlaed2).
// Inputs: m values d[] (physical order), z[] (already normalised so |z| = 1, and with
// the sign of the coupling folded in), rho > 0, ord[r] = physical index of the r-th
// smallest d.  Outputs: nd[0..k) physical indices of the non-deflated entries in
// ascending d order, df[0..m-k) deflated indices, rotations (a, b, c, s) to apply in
// order to the eigenvector columns:  (q_a, q_b) <- (c q_a + s q_b, c q_b - s q_a).
// d[] and z[] are updated in place.  Returns k.
// ---------------------------------------------------------------------------
struct DeflRot { int a, b; double c, s; };

DW_HD int deflate_scan(int m, double rho, double tol, const int* ord, double* d, double* z,
                       int* nd, int* df, DeflRot* rots, int* nrot_out) {
  int k = 0, ndf = 0, nrot = 0;
  int pj = -1;
  for (int r = 0; r < m; ++r) {
    int j = ord[r];
    if (rho * fabs(z[j]) <= tol) {
      df[ndf++] = j;
      continue;
    }
    if (pj < 0) { pj = j; continue; }
    double s = z[pj], c = z[j];
    double tau = hypot(c, s);
    double t = d[j] - d[pj];
    c /= tau;
    s = -s / tau;
    if (fabs(t * c * s) <= tol) {
      z[j] = tau;
      z[pj] = 0.0;
      rots[nrot].a = pj; rots[nrot].b = j; rots[nrot].c = c; rots[nrot].s = s;
      ++nrot;
      double dn = d[pj] * c * c + d[j] * s * s;
      d[j] = d[pj] * s * s + d[j] * c * c;
      d[pj] = dn;
      df[ndf++] = pj;
      pj = j;
    } else {
      nd[k++] = pj;
      pj = j;
    }
  }
  if (pj >= 0) nd[k++] = pj;
  *nrot_out = nrot;
  return k;
}

// ---------------------------------------------------------------------------
// Root j of the secular equation  1/rho + sum_i z_i^2 / (d_i - lambda) = 0  with
// k poles d_0 < d_1 < ... < d_{k-1} (all z_i != 0, rho > 0).  Root j lies in
// (d_j, d_{j+1}) (j < k-1) or in (d_{k-1}, d_{k-1} + rho |z|^2].
// The root is returned as lambda = d[org] + tau with org the nearer pole, so that the
// differences d_i - lambda = (d_i - d[org]) - tau are accurate to a few ulp; that is
// what the Gu/Eisenstat vector formula needs.
// Iteration: "middle way" two-pole rational interpolation, safeguarded by a bracket.
// Returns the number of iterations used (negative if the tolerance was not met).
// ---------------------------------------------------------------------------
template <class Par>
DW_HD int secular_root(int k, int j, const double* d, const double* z, double rho, const Par& par,
                       int* org_out, double* tau_out) {
  if (k == 1) {
    *org_out = 0;
    *tau_out = rho * z[0] * z[0];
    return 0;
  }
  const double rhoinv = 1.0 / rho;
  const bool last = (j == k - 1);
  const int p1 = last ? k - 2 : j;       // poles used by the rational model
  const int p2 = p1 + 1;
  int org;
  double lo, hi, tau;

  if (!last) {
    const double gap = d[j + 1] - d[j];
    const double half = 0.5 * gap;
    // f at the midpoint decides which pole is nearer
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int i = par.begin(); i < k; i += par.stride()) {
      double del = (d[i] - d[j]) - half;
      s0 += z[i] * z[i] / del;
    }
    par.sum4(s0, s1, s2, s3);
    double fmid = rhoinv + s0;
    // rational initial guess: everything except the two nearest poles frozen at the midpoint
    double zj2 = z[j] * z[j], zj12 = z[j + 1] * z[j + 1];
    double c = fmid - zj2 / (-half) - zj12 / (half);
    if (fmid > 0.0) {
      org = j; lo = 0.0; hi = half;
      // c*tau^2 - (c*gap + zj2 + zj12)*tau + zj2*gap = 0, root in (0, gap/2]
      double a = c * gap + zj2 + zj12;
      double b = zj2 * gap;
      double disc = sqrt(fabs(a * a - 4.0 * b * c));
      tau = (a > 0.0) ? 2.0 * b / (a + disc) : (a - disc) / (2.0 * c);
    } else {
      org = j + 1; lo = -half; hi = 0.0;
      // c*tau^2 - a*tau - b = 0 with a = -c*gap + zj2 + zj12, b = zj12*gap; root in [-gap/2, 0)
      double a = -c * gap + zj2 + zj12;
      double b = zj12 * gap;
      double disc = sqrt(fabs(a * a + 4.0 * b * c));
      tau = (a > 0.0) ? -2.0 * b / (a + disc) : (a - disc) / (2.0 * c);
    }
    if (!(tau > lo && tau < hi)) tau = 0.5 * (lo + hi);
  } else {
    org = k - 1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int i = par.begin(); i < k; i += par.stride()) s0 += z[i] * z[i];
    par.sum4(s0, s1, s2, s3);
    lo = 0.0;
    hi = rho * s0;
    tau = 0.5 * hi;
  }

  const double dorg = d[org];
  int it = 0;
  int status = -1;
  double best_tau = tau, best_res = 1e300;
  for (; it < 100; ++it) {
    // evaluate psi (i <= p1) and phi (i >= p2) and their derivatives at tau
    double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0;
    for (int i = par.begin(); i < k; i += par.stride()) {
      double del = (d[i] - dorg) - tau;
      double t = z[i] / del;
      double tz = t * z[i];
      if (i <= p1) { psi += tz; dpsi += t * t; }
      else         { phi += tz; dphi += t * t; }
    }
    par.sum4(psi, dpsi, phi, dphi);
    const double f = rhoinv + psi + phi;
    const double df = dpsi + dphi;
    const double erretm = 8.0 * (fabs(psi) + fabs(phi)) + 2.0 * rhoinv + fabs(tau) * df;
    const double res = fabs(f);
    if (res < best_res) { best_res = res; best_tau = tau; }
    if (res <= DW_EPS * erretm) { status = it; best_tau = tau; break; }
    if (f < 0.0) lo = tau; else hi = tau;
    if (hi - lo <= 2.0 * DW_EPS * fmax(fabs(lo), fabs(hi))) { status = it; best_tau = tau; break; }

    const double D1 = (d[p1] - dorg) - tau;
    const double D2 = (d[p2] - dorg) - tau;
    double a = (D1 + D2) * f - D1 * D2 * df;
    double b = D1 * D2 * f;
    double c = f - D1 * dpsi - D2 * dphi;
    double eta;
    if (c == 0.0) {
      eta = b / a;
    } else {
      double disc = sqrt(fabs(a * a - 4.0 * b * c));
      eta = (a <= 0.0) ? (a - disc) / (2.0 * c) : 2.0 * b / (a + disc);
    }
    if (!(f * eta < 0.0)) eta = -f / df;          // wrong direction or NaN: Newton step
    double tnew = tau + eta;
    if (!(tnew > lo && tnew < hi)) {
      // bracket safeguard: bisection, geometric where the bracket spans decades
      if (lo == 0.0) tnew = hi * 0.0625;
      else if (hi == 0.0) tnew = lo * 0.0625;
      else {
        double r = hi / lo;
        if (r > 0.0 && (r > 16.0 || r < 0.0625)) tnew = copysign(sqrt(lo * hi), lo);
        else tnew = 0.5 * (lo + hi);
      }
    }
    tau = tnew;
  }
  *org_out = org;
  *tau_out = best_tau;
  return status >= 0 ? status : -it;
}

}  // namespace dwcore
