// Host harness (TESTS ONLY) for csrc/stedc_core.h: a serial divide-and-conquer driver
// built on the same scalar routines the CUDA kernels call, so the secular solver,
// deflation scan and QL leaf solver can be checked against LAPACK on the CPU box.
// Not part of libdwhmc.so; nothing in the product path links or loads this file.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../hybrid-monte-carlo-for-d-wave-sc_b200/csrc/stedc_core.h"
#include "../../hybrid-monte-carlo-for-d-wave-sc_b200/csrc/stedc_tree.h"

using namespace dwcore;

extern "C" int host_stedc(int n, const double* d_in, const double* e_in, double* w_out, double* Z_out,
                          int leafmax, int* stats /* [0]=max secular iters, [1]=noconv, [2]=total k, [3]=total m */) {
  std::vector<double> D(d_in, d_in + n), E(std::max(n - 1, 1), 0.0);
  for (int i = 0; i + 1 < n; ++i) E[i] = e_in[i];
  std::vector<double> Za((size_t)n * n, 0.0), Zb((size_t)n * n, 0.0), S((size_t)n * n, 0.0);
  std::vector<int> perm(n);
  DcTree tree = build_dc_tree(n, leafmax);
  stats[0] = stats[1] = stats[2] = stats[3] = 0;

  // tear at every leaf boundary
  for (size_t l = 0; l + 1 < tree.leaves.size(); ++l) {
    int b = tree.leaves[l].off + tree.leaves[l].size;  // first row of the next leaf
    double r = fabs(E[b - 1]);
    D[b - 1] -= r;
    D[b] -= r;
  }
  // leaves
  for (auto& lf : tree.leaves) {
    int off = lf.off, sz = lf.size;
    std::vector<double> dd(D.begin() + off, D.begin() + off + sz), ee(sz, 0.0);
    for (int i = 0; i + 1 < sz; ++i) ee[i] = E[off + i];
    double* Z = Za.data() + (size_t)off * n + off;
    for (int i = 0; i < sz; ++i) Z[(size_t)i * n + i] = 1.0;
    auto rot = [&](int i, double c, double s) {
      for (int k = 0; k < sz; ++k) {
        double f = Z[(size_t)(i + 1) * n + k];
        Z[(size_t)(i + 1) * n + k] = s * Z[(size_t)i * n + k] + c * f;
        Z[(size_t)i * n + k] = c * Z[(size_t)i * n + k] - s * f;
      }
    };
    int info = tql_implicit(sz, dd.data(), ee.data(), rot);
    if (info) return 100 + info;
    for (int i = 0; i < sz; ++i) D[off + i] = dd[i];
    std::vector<int> idx(sz);
    for (int i = 0; i < sz; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return dd[a] < dd[b]; });
    for (int i = 0; i < sz; ++i) perm[off + i] = idx[i];
  }

  double* Zin = Za.data();
  double* Zout = Zb.data();
  for (auto& level : tree.levels) {
    for (auto& mg : level) {
      const int off = mg.off, n1 = mg.n1, m = mg.n1 + mg.n2;
      double* Q = Zin + (size_t)off * n + off;
      double* Qo = Zout + (size_t)off * n + off;
      double* Sm = S.data() + (size_t)off * n + off;
      double* d = D.data() + off;
      const double beta = E[off + n1 - 1];
      const double sgn = beta < 0 ? -1.0 : 1.0;
      double rho = 2.0 * fabs(beta);
      std::vector<double> z(m);
      const double isq2 = 1.0 / sqrt(2.0);
      for (int i = 0; i < m; ++i)
        z[i] = (i < n1 ? Q[(size_t)i * n + (n1 - 1)] : sgn * Q[(size_t)i * n + n1]) * isq2;
      // merge the two sorted orders
      std::vector<int> ord(m);
      {
        int a = 0, b = 0, r = 0;
        const int* p1 = perm.data() + off;
        const int* p2 = perm.data() + off + n1;
        while (a < n1 && b < mg.n2) {
          if (d[p1[a]] <= d[n1 + p2[b]]) ord[r++] = p1[a++];
          else ord[r++] = n1 + p2[b++];
        }
        while (a < n1) ord[r++] = p1[a++];
        while (b < mg.n2) ord[r++] = n1 + p2[b++];
      }
      double dmax = 0, zmax = 0;
      for (int i = 0; i < m; ++i) { dmax = std::max(dmax, fabs(d[i])); zmax = std::max(zmax, fabs(z[i])); }
      const double tol = 8.0 * DW_EPS * std::max(dmax, zmax);
      std::vector<int> nd(m), df(m);
      std::vector<DeflRot> rots(m);
      int nrot = 0, k = 0;
      if (rho * zmax <= tol) {
        for (int r = 0; r < m; ++r) df[r] = ord[r];
        k = 0;
      } else {
        k = deflate_scan(m, rho, tol, ord.data(), d, z.data(), nd.data(), df.data(), rots.data(), &nrot);
      }
      for (int r = 0; r < nrot; ++r) {
        double* qa = Q + (size_t)rots[r].a * n;
        double* qb = Q + (size_t)rots[r].b * n;
        const double c = rots[r].c, s = rots[r].s;
        for (int i = 0; i < m; ++i) {
          double t = c * qa[i] + s * qb[i];
          qb[i] = c * qb[i] - s * qa[i];
          qa[i] = t;
        }
      }
      stats[2] += k; stats[3] += m;
      // deflated columns go to physical positions k..m-1 of the output
      std::vector<double> dnew(m);
      for (int r = 0; r < m - k; ++r) {
        memcpy(Qo + (size_t)(k + r) * n, Q + (size_t)df[r] * n, sizeof(double) * m);
        dnew[k + r] = d[df[r]];
      }
      if (k > 0) {
        std::vector<double> dl(k), w(k);
        for (int p = 0; p < k; ++p) { dl[p] = d[nd[p]]; w[p] = z[nd[p]]; }
        SerialPar par;
        for (int j = 0; j < k; ++j) {
          int org; double tau;
          int it = secular_root(k, j, dl.data(), w.data(), rho, par, &org, &tau);
          if (it < 0) stats[1]++;
          stats[0] = std::max(stats[0], std::abs(it));
          dnew[j] = dl[org] + tau;
          for (int i = 0; i < k; ++i) {
            double del = (dl[i] - dl[org]) - tau;
            Sm[(size_t)j * n + i] = del;
          }
        }
        // Gu/Eisenstat z-hat
        std::vector<double> zh(k);
        for (int i = 0; i < k; ++i) {
          double prod = -Sm[(size_t)i * n + i];
          for (int j = 0; j < k; ++j)
            if (j != i) prod *= Sm[(size_t)j * n + i] / (dl[i] - dl[j]);
          zh[i] = copysign(sqrt(fabs(prod)), w[i]);
        }
        for (int j = 0; j < k; ++j) {
          double nrm = 0;
          for (int i = 0; i < k; ++i) {
            double v = zh[i] / Sm[(size_t)j * n + i];
            Sm[(size_t)j * n + i] = v;
            nrm += v * v;
          }
          nrm = 1.0 / sqrt(nrm);
          for (int i = 0; i < k; ++i) Sm[(size_t)j * n + i] *= nrm;
        }
        // Qo[:, j] = sum_p Q[:, nd[p]] * S[p, j]
        for (int j = 0; j < k; ++j) {
          double* out = Qo + (size_t)j * n;
          for (int i = 0; i < m; ++i) out[i] = 0.0;
          for (int p = 0; p < k; ++p) {
            const double s = Sm[(size_t)j * n + p];
            const double* q = Q + (size_t)nd[p] * n;
            for (int i = 0; i < m; ++i) out[i] += q[i] * s;
          }
        }
      }
      for (int i = 0; i < m; ++i) d[i] = dnew[i];
      std::vector<int> idx(m);
      for (int i = 0; i < m; ++i) idx[i] = i;
      std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return d[a] < d[b]; });
      for (int i = 0; i < m; ++i) perm[off + i] = idx[i];
    }
    std::swap(Zin, Zout);
  }
  for (int r = 0; r < n; ++r) {
    int c = perm[r];
    w_out[r] = D[c];
    memcpy(Z_out + (size_t)r * n, Zin + (size_t)c * n, sizeof(double) * n);
  }
  return 0;
}
