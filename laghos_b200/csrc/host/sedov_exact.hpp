// Exact Taylor - von Neumann - Sedov blast wave (uniform ambient density), the comparison solution of the reference
// driver's `-err` diagnostic (laghos.cpp:1009-1085, which uses sedov/sedov_sol.{hpp,cpp}: gamma = 1.4, rho0 = 1,
// omega = 0).  Own restatement of the published similarity solution (Sedov 1959; Kamm & Timmes, "On efficient
// generation of numerically robust Sedov solutions", LA-UR-07-2849) -- NOT a copy of the reference file:
//   * everything is written in the distance s = V - V0 from the singular end of the similarity variable, so the
//     factor c V - 1 = c s carries no cancellation;
//   * the two energy integrals (integrable end-point singularity s^(-0.68) in 3D) use tanh-sinh quadrature, whose
//     nodes cluster double-exponentially at the end points and are generated directly as distances from them;
//   * V(r) comes from a bracketed Newton iteration on log lambda.
// Checked in tests/test_sedov_exact.py against the reference's own solution compiled from its sources (test
// infrastructure, only where the reference tree exists) through the committed vectors tests/golden/sedov_exact.json.
// Host code, diagnostics only (row 8f-2).  Only the "standard" case V2 < Vs is implemented (true for every
// gamma > 1 with omega = 0 in 2D and 3D ... checked in the constructor); the singular / vacuum cases need omega > 0.
#pragma once
#include <cmath>
#include <stdexcept>

namespace lagb {

struct SedovExact
{
   int j;                         // 2: cylindrical, 3: spherical (1: planar)
   double gam, rho0, E0;
   double a, b, c, d, e;          // x1 = a V, x2 = b (c V - 1), x3 = d (1 - e V), x4 = b (1 - c V / gamma)
   double al0, al1, al2, al3, al4, al5;
   double V0, V2;                 // similarity variable: V0 at the centre, V2 at the shock
   double alpha;                  // dimensionless energy integral
   double t = 0, r2 = 0, U = 0, rho2 = 0, v2 = 0, p2 = 0;   // shock position / speed, post-shock state

   SedovExact(int dim, double gamma, double rho_ambient, double blast_energy)
      : j(dim), gam(gamma), rho0(rho_ambient), E0(blast_energy)
   {
      if (j < 1 || j > 3 || !(gam > 1.0)) { throw std::runtime_error("SedovExact: dim in 1..3 and gamma > 1"); }
      const double j2 = j + 2.0;
      a = j2*(gam + 1)/4; b = (gam + 1)/(gam - 1); c = j2*gam/2;
      e = (2 + j*(gam - 1))/2;
      d = j2*(gam + 1)/(j2*(gam + 1) - 4*e);
      al0 = 2/j2;
      al2 = -(gam - 1)/(2*(gam - 1) + j);
      al1 = j2*gam/(2*e)*(2*j*(2 - gam)/(gam*j2*j2) - al2);
      al3 = j/(2*(gam - 1) + j);
      al4 = j2*al1/(2 - gam);
      al5 = -2/(2 - gam);
      V0 = 1/c; V2 = 4/(j2*(gam + 1));
      const double Vs = 1/e;
      if (!(V2 < Vs) || gam == 2.0) { throw std::runtime_error("SedovExact: only the standard case (V2 < Vs) is implemented"); }
      alpha = energy_integral();
   }

   // log lambda(s) and its derivative with respect to s; lambda = r / r2 = x1^-al0 x2^-al2 x3^-al1
   void loglam(double s, double &ll, double &dll) const
   {
      const double V = V0 + s;
      ll = -al0*std::log(a*V) - al2*std::log(b*c*s) - al1*std::log(d*(1 - e*V));
      dll = -al0/V - al2/s + al1*e/(1 - e*V);
   }
   // similarity profiles at s: f = v / v2, g = rho / rho2, h = p / p2
   void profiles(double s, double lam, double &f, double &g, double &h) const
   {
      const double V = V0 + s;
      const double x1 = a*V, x2 = b*c*s, x3 = d*(1 - e*V), x4 = b*(1 - c*V/gam);
      f = x1*lam;
      g = std::pow(x2, al3)*std::pow(x3, al4)*std::pow(x4, al5);
      h = std::pow(x1, al0*j)*std::pow(x3, al4 - 2*al1)*std::pow(x4, 1 + al5);
   }

   // alpha = geom (2^(j-2) J1 + 2^(j-1)/(gamma-1) J2), geom = pi for j > 1 (Kamm & Timmes eqs. 55-57 with omega = 0):
   // kinetic and internal energy of the similarity profiles as integrals over V, here over s = V - V0 in (0, V2 - V0):
   //   dJ1 = (gamma+1)/(gamma-1) V^2 lambda^(j+2) g dloglam,
   //   dJ2 = (gamma+1)/(2 gamma) V^2 (gamma - c V)/(c s) lambda^(j+2) g dloglam      (1 - c V = -c s)
   double energy_integral() const
   {
      const double L = V2 - V0;
      auto integrand = [&](double s, double &k1, double &k2)
      {
         double ll, dll, f, g, h;
         loglam(s, ll, dll);
         const double V = V0 + s, lam = std::exp(ll);
         profiles(s, lam, f, g, h);
         const double common = V*V*std::exp((j + 2)*ll)*g*dll;
         k1 = (gam + 1)/(gam - 1)*common;
         k2 = (gam + 1)/(2*gam)*(gam - c*V)/(c*s)*common;
      };
      // tanh-sinh: s = L/(1 + exp(-2u)), u = (pi/2) sinh(tau); the distance to the nearer end is formed without
      // cancellation (the singular end s -> 0 is the one that needs it)
      const double hstep = 1.0/64;
      double J1 = 0, J2 = 0;
      for (int k = -512; k <= 512; k++)
      {
         const double tau = k*hstep, u = 0.5*M_PI*std::sinh(tau);
         const double em = std::exp(-2*std::fabs(u));
         const double near = L*em/(1 + em);                            // distance to the nearer end
         if (!(near > 1e-300)) { continue; }
         const double w = L*M_PI*std::cosh(tau)*em/((1 + em)*(1 + em))*hstep;   // (ds/du) (du/dtau) dtau
         double k1, k2;
         integrand((u < 0) ? near : L - near, k1, k2);
         if (std::isfinite(k1) && std::isfinite(k2)) { J1 += w*k1; J2 += w*k2; }
      }
      const double geom = (j == 1) ? 1.0 : M_PI;
      return geom*(std::pow(2.0, j - 2)*J1 + std::pow(2.0, j - 1)/(gam - 1)*J2);
   }

   // replace the energy integral (tests: the reference's quadrature of it stops ~1e-5 short, see the header of
   // tests/test_sedov_exact.py; with its value the profiles agree to round-off)
   void set_alpha(double al) { alpha = al; if (t > 0) { set_time(t); } }

   void set_time(double time)
   {
      t = time;
      const double j2 = j + 2.0;
      r2 = std::pow(E0/(alpha*rho0), 1/j2)*std::pow(t, 2/j2);
      U = 2/j2*r2/t;
      rho2 = b*rho0; v2 = 2/(gam + 1)*U; p2 = 2/(gam + 1)*rho0*U*U;
   }

   void eval(double r, double &rho, double &v, double &p) const
   {
      if (r >= r2) { rho = rho0; v = 0; p = 0; return; }
      const double L = V2 - V0;
      if (!(r > 0)) { rho = 0; v = 0; p = p2*h_centre(); return; }
      const double target = std::log(r/r2);
      // bracketed Newton on log lambda(s) = log(r / r2), s in (0, L): log lambda increases with s, ~ -al2 log s at s -> 0
      double lo = 0, hi = L, s = L*std::pow(r/r2, -1/al2);
      if (!(s > 0 && s < L)) { s = 0.5*L; }
      for (int it = 0; it < 100; it++)
      {
         double ll, dll;
         loglam(s, ll, dll);
         const double res = ll - target;
         if (res > 0) { hi = s; } else { lo = s; }
         double sn = s - res/dll;
         if (!(sn > lo && sn < hi)) { sn = (lo > 0) ? std::sqrt(lo*hi) : 0.01*hi; }
         const bool done = std::fabs(sn - s) <= 1e-15*s;
         s = sn;
         if (done) { break; }
      }
      double ll, dll, f, g, h;
      loglam(s, ll, dll);
      profiles(s, std::exp(ll), f, g, h);
      rho = rho2*g; v = v2*f; p = p2*h;
   }
   // pressure ratio at the centre: h(s -> 0) is finite (x1^(al0 j) x3^(..) x4^(1+al5) at V0)
   double h_centre() const
   {
      const double x1 = a*V0, x3 = d*(1 - e*V0), x4 = b*(1 - c*V0/gam);
      return std::pow(x1, al0*j)*std::pow(x3, al4 - 2*al1)*std::pow(x4, 1 + al5);
   }
};

} // namespace lagb
