// TEST INFRASTRUCTURE (oracle/): C wrapper around the REFERENCE's own Taylor-von Neumann-Sedov solution
// (/root/reference/sedov/sedov_sol.{hpp,cpp}, used by the driver's -err diagnostic, laghos.cpp:1009-1085).
// The reference sources are compiled where they lie (no copy in this repo): `make oracle/_ref/libsedov_ref.so`
// when /root/reference exists; the output lives in oracle/_ref/ (git-ignored).  Only tests/ and
// tools/make_sedov_golden.py load it, to pin the product's own restatement (laghos_b200/csrc/host/sedov_exact.hpp).
#include "sedov_sol.hpp"

extern "C" {

// info: alpha, r2, U, rho2, v2, p2
int sedov_ref_eval(int dim, double gamma, double rho0, double blast_energy, double omega, double t,
                   int n, const double *r, double *rho, double *v, double *P, double *info)
{
   try
   {
      SedovSol s(dim, gamma, rho0, blast_energy, omega);
      s.SetTime(t);
      for (int i = 0; i < n; i++) { s.EvalSol(r[i], rho[i], v[i], P[i]); }
      if (info) { info[0] = s.alpha; info[1] = s.r2; info[2] = s.U; info[3] = s.rho2; info[4] = s.v2; info[5] = s.p2; }
      return 0;
   }
   catch (...) { return 1; }
}

}
