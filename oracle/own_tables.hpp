// TEST INFRASTRUCTURE — the oracle's OWN 1D tables, quadrature weights and gather map.
//
// The oracle takes meshes and initial conditions from the product's host set-up (problem.hpp: a specification
// table), but the operator-side data -- basis tables B, G, BL, the quadrature weights and the element restriction --
// is rebuilt here with algorithms that share nothing with laghos_b200/csrc/host/fe_tables.hpp:
//   * Gauss-Legendre points / weights: Golub-Welsch (eigenvalues of the Legendre Jacobi matrix by the implicit QL
//     method, weights from the first eigenvector components), not Newton on P_n;
//   * Gauss-Lobatto nodes: end points + the Gauss-Jacobi(1,1) points (roots of P'_p), again by Golub-Welsch;
//   * Lagrange values / derivatives: barycentric formulas, not products of differences;
//   * Bernstein values: the closed form C(p,l) x^l (1-x)^(p-l), not the de Casteljau recurrence;
//   * gather map: MFEM's lexicographic ElementRestriction of a Cartesian block, written from its definition.
// install_own_tables() first CHECKS the product's tables against these (1e-13; a disagreement throws, so every oracle
// run is also a test of fe_tables.hpp) and then replaces them in the oracle's private Problem instance: a bug in
// the product's tables or map can no longer cancel in a GPU-vs-oracle comparison (reference behaviours restated:
// SURVEY App. B.1; laghos_solver.cpp:145-147 for the rule, laghos.cpp:494-495 for the bases).
#pragma once
#include "../laghos_b200/csrc/host/problem.hpp"
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

// eigenvalues d (in: diagonal) of the symmetric tridiagonal matrix with sub-diagonal e[0..n-2], and the first row z
// of its orthonormal eigenvector matrix (in: z = e_1): implicit QL with Wilkinson shifts
inline void tridiag_ql(std::vector<long double> &d, std::vector<long double> &e, std::vector<long double> &z)
{
   const int n = (int)d.size();
   e.resize(n, 0.0L);
   const long double eps = 1e-19L;
   for (int l = 0; l < n; l++)
   {
      int iter = 0, m;
      do
      {
         for (m = l; m < n - 1; m++)
         {
            const long double dd = fabsl(d[m]) + fabsl(d[m + 1]);
            if (fabsl(e[m]) <= eps*dd) { break; }
         }
         if (m != l)
         {
            if (++iter > 300) { throw std::runtime_error("tridiag_ql: no convergence"); }
            long double g = (d[l + 1] - d[l])/(2.0L*e[l]);
            long double r = hypotl(g, 1.0L);
            g = d[m] - d[l] + e[l]/(g + copysignl(r, g));
            long double s = 1.0L, c = 1.0L, p = 0.0L;
            int i;
            for (i = m - 1; i >= l; i--)
            {
               long double f = s*e[i];
               const long double b = c*e[i];
               r = hypotl(f, g);
               e[i + 1] = r;
               if (r == 0.0L) { d[i + 1] -= p; e[m] = 0.0L; break; }
               s = f/r; c = g/r;
               g = d[i + 1] - p;
               r = (d[i] - g)*s + 2.0L*c*b;
               p = s*r;
               d[i + 1] = g + p;
               g = c*r - b;
               f = z[i + 1];
               z[i + 1] = s*z[i] + c*f;
               z[i] = c*z[i] - s*f;
            }
            if (r == 0.0L && i >= l) { continue; }
            d[l] -= p; e[l] = g; e[m] = 0.0L;
         }
      }
      while (m != l);
   }
}

// Gauss rule for the weight (1-x)^a (1+x)^a on [-1,1], a = 0 (Legendre) or 1 (Jacobi(1,1)), mapped to [0,1], ascending
inline void golub_welsch_01(int n, int a, std::vector<double> &x, std::vector<double> &w)
{
   std::vector<long double> d(n, 0.0L), e(n, 0.0L), z(n, 0.0L);
   for (int k = 1; k < n; k++)
   {
      const long double kk = k;
      e[k - 1] = (a == 0) ? kk/sqrtl(4.0L*kk*kk - 1.0L) : sqrtl(kk*(kk + 2.0L)/((2.0L*kk + 1.0L)*(2.0L*kk + 3.0L)));
   }
   z[0] = 1.0L;
   tridiag_ql(d, e, z);
   const long double mu0 = (a == 0) ? 2.0L : 4.0L/3.0L;       // integral of the weight
   std::vector<int> idx(n);
   for (int i = 0; i < n; i++) { idx[i] = i; }
   for (int i = 1; i < n; i++) { for (int j = i; j > 0 && d[idx[j]] < d[idx[j - 1]]; j--) { std::swap(idx[j], idx[j - 1]); } }
   x.resize(n); w.resize(n);
   for (int i = 0; i < n; i++)
   {
      x[i] = (double)((d[idx[i]] + 1.0L)/2.0L);
      w[i] = (double)(mu0*z[idx[i]]*z[idx[i]]/2.0L);
   }
}

// barycentric Lagrange basis on nodes xi at x: values and derivatives
inline void barycentric(const std::vector<double> &xi, double x, double *val, double *der)
{
   const int n = (int)xi.size();
   std::vector<long double> bw(n, 1.0L);
   for (int j = 0; j < n; j++) { for (int m = 0; m < n; m++) { if (m != j) { bw[j] /= ((long double)xi[j] - xi[m]); } } }
   int hit = -1;
   for (int j = 0; j < n; j++) { if (x == xi[j]) { hit = j; } }
   if (hit < 0)
   {
      long double den = 0.0L;
      for (int k = 0; k < n; k++) { den += bw[k]/((long double)x - xi[k]); }
      for (int j = 0; j < n; j++)
      {
         const long double lj = bw[j]/((long double)x - xi[j])/den;
         long double s = 0.0L;                                   // l_j'(x) = l_j(x) sum_{m != j} 1/(x - x_m)
         for (int m = 0; m < n; m++) { if (m != j) { s += 1.0L/((long double)x - xi[m]); } }
         val[j] = (double)lj; der[j] = (double)(lj*s);
      }
   }
   else
   {
      long double self = 0.0L;
      for (int j = 0; j < n; j++)
      {
         val[j] = (j == hit) ? 1.0 : 0.0;
         if (j != hit)
         {
            const long double dj = (bw[j]/bw[hit])/((long double)xi[hit] - xi[j]);
            der[j] = (double)dj; self -= dj;
         }
      }
      der[hit] = (double)self;
   }
}

inline double binomial(int n, int k) { double b = 1.0; for (int i = 1; i <= k; i++) { b = b*(n - k + i)/i; } return b; }

inline void install_own_tables(lagb::Problem &P)
{
   const int ok = P.spec.ok, ot = P.spec.ot, D = ok + 1, L = ot + 1;
   const int oq = (P.spec.oq > 0) ? P.spec.oq : 3*ok + ot - 1;     // laghos_solver.cpp:145-147
   const int Q = (oq | 1)/2 + 1;                                  // IntRules.Get(SEGMENT, oq)
   if (D != P.D1D || L != P.L1D || Q != P.Q1D) { throw std::runtime_error("oracle tables: sizes disagree with the host set-up"); }
   std::vector<double> qx, qw, gll(D), tmp;
   golub_welsch_01(Q, 0, qx, qw);
   gll[0] = 0.0; gll[D - 1] = 1.0;
   if (D > 2) { std::vector<double> in, inw; golub_welsch_01(D - 2, 1, in, inw); for (int i = 0; i < D - 2; i++) { gll[i + 1] = in[i]; } }
   std::vector<double> B((size_t)Q*D), G((size_t)Q*D), Bt((size_t)Q*D), Gt((size_t)Q*D), BL((size_t)Q*L), BLt((size_t)Q*L);
   std::vector<double> v(D), d(D);
   for (int q = 0; q < Q; q++)
   {
      barycentric(gll, qx[q], v.data(), d.data());
      for (int i = 0; i < D; i++) { B[q + Q*i] = Bt[i + D*q] = v[i]; G[q + Q*i] = Gt[i + D*q] = d[i]; }
      for (int l = 0; l < L; l++)
      {
         const double b = binomial(ot, l)*std::pow(qx[q], l)*std::pow(1.0 - qx[q], ot - l);
         BL[q + Q*l] = BLt[l + L*q] = b;
      }
   }
   std::vector<double> qweights(P.NQ);
   for (int q = 0; q < P.NQ; q++)
   {
      const int qx_ = q % Q, qy_ = (q/Q) % Q, qz_ = q/(Q*Q);
      qweights[q] = qw[qx_]*qw[qy_]*((P.dim == 3) ? qw[qz_] : 1.0);
   }
   // lexicographic element restriction of the rank's Cartesian block: local node (kx,ky,kz) of element (ix,iy,iz) is
   // lattice point (ix p + kx, iy p + ky, iz p + kz) of the (n p + 1)^dim H1 lattice, x fastest
   std::vector<int> map((size_t)P.NE*P.ND);
   const int nx = P.nloc[0], ny = P.nloc[1], Lx = nx*ok + 1, Ly = ny*ok + 1;
   for (int e = 0; e < P.NE; e++)
   {
      const int iz = e/(nx*ny), iy = (e - iz*nx*ny)/nx, ix = e - iz*nx*ny - iy*nx;
      for (int i = 0; i < P.ND; i++)
      {
         const int kz = i/(D*D), ky = (i - kz*D*D)/D, kx = i - kz*D*D - ky*D;
         map[(size_t)e*P.ND + i] = (ix*ok + kx) + Lx*((iy*ok + ky) + Ly*(iz*ok + kz));
      }
   }
   // cross-check of the product's host set-up, then replace
   auto check = [](const std::vector<double> &a, const std::vector<double> &b, const char *what)
   {
      if (a.size() != b.size()) { throw std::runtime_error(std::string("oracle tables: size of ") + what); }
      for (size_t i = 0; i < a.size(); i++)
      {
         if (!(std::fabs(a[i] - b[i]) <= 1e-13*(1.0 + std::fabs(a[i])))) { throw std::runtime_error(std::string("oracle tables: the host set-up disagrees on ") + what); }
      }
   };
   check(P.tab.qx, qx, "Gauss points"); check(P.tab.qw, qw, "Gauss weights"); check(P.tab.gll, gll, "Gauss-Lobatto nodes");
   check(P.tab.B, B, "B"); check(P.tab.G, G, "G"); check(P.tab.BL, BL, "BL"); check(P.qweights, qweights, "quadrature weights");
   if (P.h1_map != map) { throw std::runtime_error("oracle tables: the host set-up disagrees on the gather map"); }
   P.tab.qx = qx; P.tab.qw = qw; P.tab.B = B; P.tab.G = G; P.tab.Bt = Bt; P.tab.Gt = Gt; P.tab.BL = BL; P.tab.BLt = BLt;
   P.qweights = qweights; P.h1_map = map;
}

} // namespace oracle
