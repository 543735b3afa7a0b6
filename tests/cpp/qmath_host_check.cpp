// Host check of the device point-physics math (laghos_b200/csrc/device/qmath.cuh, host code path of the
// same functions) against the oracle's restatement of MFEM's closed-form routines (oracle/smallmat.hpp).
// Prints "name max_err" lines; tests/test_qmath_host.py asserts the thresholds.
#include "../../laghos_b200/csrc/device/qmath.cuh"
#include "../../oracle/smallmat.hpp"
#include <cstdio>
#include <random>
#include <algorithm>

using namespace lagb;
namespace sm = oracle::sm;

int main()
{
   std::mt19937_64 rng(12345);
   std::uniform_real_distribution<double> U(-1.0, 1.0);
   // 1. cos(acos(x)/3)
   double e_cos = 0.0;
   for (int i = 0; i <= 2000000; i++)
   {
      const double x = -0.9 + 1.9*i/2000000.0;
      const long double ref = cosl(acosl((long double)x)/3.0L);
      e_cos = std::max(e_cos, (double)fabsl((long double)qm::cos_acos_third(x) - ref));
   }
   printf("cos_acos_third %.3e\n", e_cos);
   // 2. symmetric eigenproblem: random, clustered (double / near-double / near-triple) spectra
   double e_val = 0.0, e_vec = 0.0; long n_vec = 0;
   double e_sv = 0.0;
   for (int trial = 0; trial < 400000; trial++)
   {
      // random orthogonal basis from a random matrix (Gram-Schmidt)
      double q[3][3];
      for (auto &r : q) { for (double &v : r) { v = U(rng); } }
      auto dot = [](const double *a, const double *b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; };
      auto nrm = [&](double *a) { const double n = sqrt(dot(a, a)); a[0] /= n; a[1] /= n; a[2] /= n; };
      nrm(q[0]);
      { const double d = dot(q[1], q[0]); for (int k = 0; k < 3; k++) { q[1][k] -= d*q[0][k]; } nrm(q[1]); }
      q[2][0] = q[0][1]*q[1][2] - q[0][2]*q[1][1]; q[2][1] = q[0][2]*q[1][0] - q[0][0]*q[1][2]; q[2][2] = q[0][0]*q[1][1] - q[0][1]*q[1][0];
      double lam[3];
      const int kind = trial % 8;
      const double scale = pow(10.0, 3.0*U(rng));
      lam[0] = U(rng); lam[1] = U(rng); lam[2] = U(rng);
      if (kind == 1) { lam[1] = lam[0]; }                                  // exact double
      if (kind == 2) { lam[1] = lam[0]*(1.0 + 1e-9*U(rng)); }              // near double
      if (kind == 3) { lam[1] = lam[0]*(1.0 + 1e-5*U(rng)); lam[2] = lam[0]*(1.0 + 1e-5*U(rng)); }   // near triple
      if (kind == 4) { lam[2] = lam[1] = lam[0] + 0.3; }                   // smallest single, double above
      if (kind == 5) { lam[0] = -fabs(lam[0]); lam[1] = lam[2] = fabs(lam[1]); }
      double A[9];
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
      {
         double a = 0.0; for (int k = 0; k < 3; k++) { a += scale*lam[k]*q[k][i]*q[k][j]; }
         A[i + 3*j] = a;
      }
      for (int i = 0; i < 3; i++) for (int j = 0; j < i; j++) { A[i + 3*j] = A[j + 3*i]; }
      double l_ref[3], v_ref[9];
      sm::CalcEigenvalues<3>(A, l_ref, v_ref);
      double mu, x0, x1, x2;
      qm::min_eig3(A[0], A[3], A[6], A[4], A[7], A[8], mu, x0, x1, x2);
      const double an = scale*std::max({fabs(lam[0]), fabs(lam[1]), fabs(lam[2])});
      e_val = std::max(e_val, fabs(mu - l_ref[0])/an);
      // the eigenvector is only defined when the smallest eigenvalue is separated
      const double gap = (l_ref[1] - l_ref[0])/an;
      if (gap > 1e-3)
      {
         const double c = fabs(x0*v_ref[0] + x1*v_ref[1] + x2*v_ref[2]);
         e_vec = std::max(e_vec, fabs(1.0 - c)); n_vec++;
      }
      // 3. smallest singular value of J = Q1 diag(s) Q2^t-like matrices: use J = sym part + perturbation
      double J[9];
      const int kj = trial % 5;
      for (int k = 0; k < 9; k++) { J[k] = (k % 4 == 0 ? 1.0 : 0.0) + ((kj == 0) ? 1e-16 : (kj == 1) ? 1e-8 : (kj == 2) ? 1e-3 : 0.3)*U(rng); }
      if (kj == 4) { for (int k = 0; k < 9; k++) { J[k] = A[k]/an + (k % 4 == 0 ? 1.5 : 0.0); } }
      for (double &v : J) { v *= scale; }
      const double s_ref = sm::CalcSingularvalue<3>(J, 2);
      const double s_new = qm::min_sv3(J[0], J[1], J[2], J[3], J[4], J[5], J[6], J[7], J[8]);
      e_sv = std::max(e_sv, fabs(s_new - s_ref)/s_ref);
   }
   printf("min_eig3_value %.3e\n", e_val);
   printf("min_eig3_vector %.3e %ld\n", e_vec, n_vec);
   printf("min_sv3 %.3e\n", e_sv);
   // 4. zero matrix
   { double mu, x0, x1, x2; qm::min_eig3(0, 0, 0, 0, 0, 0, mu, x0, x1, x2); printf("zero %g %g %g %g\n", mu, x0, x1, x2); }
   return 0;
}
