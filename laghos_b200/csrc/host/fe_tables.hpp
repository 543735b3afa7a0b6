// 1D quadrature and basis tables for the partial-assembly hot path.
//
// What the reference gets from MFEM (not in /root/reference):
//   * tensor Gauss-Legendre rule with Q1D = (oq|1)/2 + 1 points on [0,1]
//     (IntRules.Get(CUBE, 3*ok+ot-1), reference laghos_solver.cpp:145-147);
//   * H1 basis = Lagrange polynomials at the Gauss-Lobatto points (H1_FECollection,
//     reference laghos.cpp:495), tables B(q,d), G(q,d) = DofToQuad::B/G used by
//     reference laghos_assembly.cpp:141-142, 561-563;
//   * L2 basis = Bernstein polynomials (BasisType::Positive, laghos.cpp:494).
// In-tree corroboration of these choices: amr/laghos_assembly.cpp:32-58.
//
// Layouts follow the reference kernels: B[q + Q1D*d] ("B(q,d)"),
// Bt[d + D1D*q] (reference laghos_assembly.cpp:153-155, 304-306).
#pragma once
#include <cmath>
#include <vector>
#include <stdexcept>

namespace lagb {

// Legendre polynomial P_n(z) and derivative on [-1,1] (long double for tables).
inline void legendre(int n, long double z, long double &p, long double &dp)
{
   long double p0 = 1.0L, p1 = z;
   if (n == 0) { p = 1.0L; dp = 0.0L; return; }
   for (int k = 2; k <= n; k++)
   {
      const long double pk = ((2*k - 1)*z*p1 - (k - 1)*p0)/k;
      p0 = p1; p1 = pk;
   }
   p = p1;
   dp = n*(z*p1 - p0)/(z*z - 1.0L);
}

// n-point Gauss-Legendre rule mapped to [0,1], points ascending.
inline void gauss_legendre_01(int n, std::vector<double> &x, std::vector<double> &w)
{
   x.assign(n, 0.0); w.assign(n, 0.0);
   const long double pi = 3.14159265358979323846264338327950288L;
   for (int i = 0; i < (n + 1)/2; i++)
   {
      long double z = cosl(pi*(i + 0.75L)/(n + 0.5L));
      for (int it = 0; it < 100; it++)
      {
         long double p, dp; legendre(n, z, p, dp);
         const long double dz = p/dp;
         z -= dz;
         if (fabsl(dz) < 1e-19L) { break; }
      }
      long double p, dp; legendre(n, z, p, dp);
      const long double wt = 2.0L/((1.0L - z*z)*dp*dp);
      // z is the i-th largest root
      x[n-1-i] = (double)((1.0L + z)/2.0L);
      x[i]     = (double)((1.0L - z)/2.0L);
      w[n-1-i] = w[i] = (double)(wt/2.0L);
   }
}

// n-point Gauss-Lobatto points on [0,1] (n >= 2), ascending.
inline void gauss_lobatto_01(int n, std::vector<double> &x)
{
   x.assign(n, 0.0);
   if (n < 2) { throw std::runtime_error("gauss_lobatto_01: n < 2"); }
   x[0] = 0.0; x[n-1] = 1.0;
   const int m = n - 1; // interior points are the roots of P'_m
   const long double pi = 3.14159265358979323846264338327950288L;
   for (int i = 1; i <= (n - 2 + 1)/2; i++)
   {
      // initial guess: Chebyshev-Gauss-Lobatto
      long double z = cosl(pi*i/m);
      for (int it = 0; it < 100; it++)
      {
         long double p, dp; legendre(m, z, p, dp);
         // P''_m from the Legendre ODE: (1-z^2) P'' = 2 z P' - m(m+1) P
         const long double ddp = (2.0L*z*dp - m*(m + 1.0L)*p)/(1.0L - z*z);
         const long double dz = dp/ddp;
         z -= dz;
         if (fabsl(dz) < 1e-19L) { break; }
      }
      x[n-1-i] = (double)((1.0L + z)/2.0L);
      x[i]     = (double)((1.0L - z)/2.0L);
   }
   if (n % 2 == 1) { x[n/2] = 0.5; }
}

// Lagrange basis on nodes xi[0..n) evaluated at x: values and derivatives.
inline void lagrange_eval(const std::vector<double> &xi, double x,
                          double *val, double *der)
{
   const int n = (int)xi.size();
   for (int j = 0; j < n; j++)
   {
      long double v = 1.0L, d = 0.0L;
      for (int m = 0; m < n; m++)
      {
         if (m == j) { continue; }
         v *= ((long double)x - xi[m])/((long double)xi[j] - xi[m]);
      }
      for (int k = 0; k < n; k++)
      {
         if (k == j) { continue; }
         long double t = 1.0L/((long double)xi[j] - xi[k]);
         for (int m = 0; m < n; m++)
         {
            if (m == j || m == k) { continue; }
            t *= ((long double)x - xi[m])/((long double)xi[j] - xi[m]);
         }
         d += t;
      }
      val[j] = (double)v;
      if (der) { der[j] = (double)d; }
   }
}

// Bernstein basis of degree p at x: C(p,l) x^l (1-x)^(p-l), l = 0..p.
inline void bernstein_eval(int p, double x, double *val)
{
   // de Casteljau-style recurrence (all terms non-negative on [0,1]).
   val[0] = 1.0;
   for (int n = 1; n <= p; n++)
   {
      val[n] = val[n-1]*x;
      for (int l = n - 1; l > 0; l--)
      {
         val[l] = val[l]*(1.0 - x) + val[l-1]*x;
      }
      val[0] *= (1.0 - x);
   }
}

struct Tables1D
{
   int D1D = 0, L1D = 0, Q1D = 0;
   std::vector<double> qx, qw;      // Gauss-Legendre points / weights on [0,1]
   std::vector<double> gll;         // H1 nodes (Gauss-Lobatto), D1D
   std::vector<double> gl_l2;       // nodal-L2 nodes (Gauss-Legendre), L1D
   std::vector<double> B, G;        // H1: B(q,d), G(q,d)  -> [q + Q1D*d]
   std::vector<double> Bt, Gt;      // H1: Bt(d,q)         -> [d + D1D*q]
   std::vector<double> BL, BLt;     // L2 Bernstein: BL(q,l) -> [q + Q1D*l], BLt(l,q)
   // L2 nodal(Gauss-Legendre) -> Bernstein change of basis, 1D: c = N2B * f,
   // N2B[l + L1D*i] (Bernstein coeff l from nodal value i).
   std::vector<double> N2B;

   // order_q <= 0 means the reference default 3*ok + ot - 1
   // (reference laghos_solver.cpp:146).
   void build(int ok, int ot, int order_q)
   {
      D1D = ok + 1; L1D = ot + 1;
      const int oq = (order_q > 0) ? order_q : 3*ok + ot - 1;
      Q1D = (oq | 1)/2 + 1;
      gauss_legendre_01(Q1D, qx, qw);
      gauss_lobatto_01(D1D, gll);
      std::vector<double> wtmp;
      gauss_legendre_01(L1D, gl_l2, wtmp);
      B.assign(Q1D*D1D, 0.0); G.assign(Q1D*D1D, 0.0);
      Bt.assign(Q1D*D1D, 0.0); Gt.assign(Q1D*D1D, 0.0);
      BL.assign(Q1D*L1D, 0.0); BLt.assign(Q1D*L1D, 0.0);
      std::vector<double> v(D1D), d(D1D), bl(L1D);
      for (int q = 0; q < Q1D; q++)
      {
         lagrange_eval(gll, qx[q], v.data(), d.data());
         for (int i = 0; i < D1D; i++)
         {
            B[q + Q1D*i] = v[i];  Bt[i + D1D*q] = v[i];
            G[q + Q1D*i] = d[i];  Gt[i + D1D*q] = d[i];
         }
         bernstein_eval(ot, qx[q], bl.data());
         for (int l = 0; l < L1D; l++)
         {
            BL[q + Q1D*l] = bl[l]; BLt[l + L1D*q] = bl[l];
         }
      }
      // Invert the 1D Bernstein Vandermonde at the nodal-L2 points: the
      // reference projects nodal L2 -> positive L2 with
      // GridFunction::ProjectGridFunction (laghos.cpp:595, 622), which for equal
      // dof counts is the exact change of basis (MFEM, not in tree).
      std::vector<long double> V(L1D*L1D), I(L1D*L1D, 0.0L);
      for (int i = 0; i < L1D; i++)
      {
         bernstein_eval(ot, gl_l2[i], bl.data());
         for (int l = 0; l < L1D; l++) { V[i + L1D*l] = bl[l]; }
         I[i + L1D*i] = 1.0L;
      }
      // Gauss-Jordan with partial pivoting on V (column-major V[i + L1D*l]).
      for (int c = 0; c < L1D; c++)
      {
         int piv = c;
         for (int r = c + 1; r < L1D; r++)
         {
            if (fabsl(V[r + L1D*c]) > fabsl(V[piv + L1D*c])) { piv = r; }
         }
         for (int k = 0; k < L1D; k++)
         {
            std::swap(V[c + L1D*k], V[piv + L1D*k]);
            std::swap(I[c + L1D*k], I[piv + L1D*k]);
         }
         const long double s = 1.0L/V[c + L1D*c];
         for (int k = 0; k < L1D; k++) { V[c + L1D*k] *= s; I[c + L1D*k] *= s; }
         for (int r = 0; r < L1D; r++)
         {
            if (r == c) { continue; }
            const long double f = V[r + L1D*c];
            for (int k = 0; k < L1D; k++)
            {
               V[r + L1D*k] -= f*V[c + L1D*k];
               I[r + L1D*k] -= f*I[c + L1D*k];
            }
         }
      }
      N2B.assign(L1D*L1D, 0.0);
      for (int l = 0; l < L1D; l++)
         for (int i = 0; i < L1D; i++) { N2B[l + L1D*i] = (double)I[l + L1D*i]; }
   }
};

} // namespace lagb
