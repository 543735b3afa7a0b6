// End-of-run velocity error norms of the reference driver (laghos.cpp:970-982): for problems 0 and 4 the exact
// velocity is constant in time, and the driver prints
//    L_inf = v_gf.ComputeMaxError(v_coeff), L_1 = v_gf.ComputeL1Error(v_coeff), L_2 = v_gf.ComputeL2Error(v_coeff).
// MFEM (not in the reference tree; restated from its published algorithm, GridFunction::ComputeLpError /
// ComputeL2Error with a VectorCoefficient): per element the tensor Gauss-Legendre rule of order 2 p + 3
// (p = the H1 order: p + 2 points per axis), at every point the Euclidean norm err = |v_h - v0(x_h)| with the exact
// field evaluated at the CURRENT physical position x_h of the point, and
//    L_inf = max err,   L_1 = sum w detJ err,   L_2 = sqrt(sum w detJ err^2).
// Host code on the host copy of the state, off the timed path (diagnostics row 8f-2).  A partitioned run combines
// the rank values with max / sum / sum of squares (ParGridFunction does the same reductions).
#pragma once
#include "problem.hpp"
#include "sedov_exact.hpp"
#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

namespace lagb {

// element ranges over the host threads (diagnostics at benchmark size: 262 144 elements x 9^3 points); the partial
// results are combined in chunk order, so the value is reproducible for a given thread count
template <class F> inline void for_element_chunks(int NE, int &nchunks, F &&body)
{
   const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
   nchunks = std::max(1, std::min(std::min(hw, 256), NE/64));
   std::vector<std::thread> pool;
   for (int k = 0; k < nchunks; k++)
   {
      const int e0 = (int)((long long)NE*k/nchunks), e1 = (int)((long long)NE*(k + 1)/nchunks);
      if (nchunks == 1) { body(k, e0, e1); } else { pool.emplace_back([&body, k, e0, e1]() { body(k, e0, e1); }); }
   }
   for (auto &t : pool) { t.join(); }
}

// out[0] = max, out[1] = L1 sum, out[2] = L2 sum of SQUARES (take the root after the rank reduction)
inline void velocity_error_sums(const Problem &P, const double *S, double out[3])
{
   const int dim = P.dim, D = P.D1D, p = P.spec.ok, n = p + 2;    // IntRules.Get(geom, 2 p + 3): (2p+3)/2 + 1 points
   std::vector<double> gx, gw;
   gauss_legendre_01(n, gx, gw);
   std::vector<double> B((size_t)n*D), G((size_t)n*D);
   for (int q = 0; q < n; q++) { lagrange_eval(P.tab.gll, gx[q], &B[(size_t)q*D], &G[(size_t)q*D]); }
   const int64_t nd = P.ndofs_h1;
   const double *X = S, *V = S + P.h1_vsize();
   ICs ic {P.spec.problem, dim};
   const int nz = (dim == 3) ? n : 1, DZ = (dim == 3) ? D : 1;
   std::vector<double> part(3*256, 0.0);
   int nchunks = 1;
   for_element_chunks(P.NE, nchunks, [&](int chunk, int e_begin, int e_end)
   {
   double emax = 0.0, e1 = 0.0, e2 = 0.0;
   for (int e = e_begin; e < e_end; e++)
   {
      const int *map = &P.h1_map[(size_t)e*P.ND];
      for (int qz = 0; qz < nz; qz++)
         for (int qy = 0; qy < n; qy++)
            for (int qx = 0; qx < n; qx++)
            {
               double x[3] = {0, 0, 0}, v[3] = {0, 0, 0}, J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
               for (int kz = 0; kz < DZ; kz++)
                  for (int ky = 0; ky < D; ky++)
                     for (int kx = 0; kx < D; kx++)
                     {
                        const double bx = B[(size_t)qx*D + kx], by = B[(size_t)qy*D + ky];
                        const double bz = (dim == 3) ? B[(size_t)qz*D + kz] : 1.0;
                        const double g[3] = {G[(size_t)qx*D + kx]*by*bz, bx*G[(size_t)qy*D + ky]*bz,
                                             (dim == 3) ? bx*by*G[(size_t)qz*D + kz] : 0.0};
                        const double b = bx*by*bz;
                        const int64_t id = map[kx + D*(ky + D*kz)];
                        for (int c = 0; c < dim; c++)
                        {
                           const double xc = X[(size_t)c*nd + id];
                           x[c] += b*xc; v[c] += b*V[(size_t)c*nd + id];
                           for (int d = 0; d < dim; d++) { J[c][d] += g[d]*xc; }
                        }
                     }
               const double det = (dim == 2) ? J[0][0]*J[1][1] - J[0][1]*J[1][0]
                                  : J[0][0]*(J[1][1]*J[2][2] - J[1][2]*J[2][1])
                                  - J[0][1]*(J[1][0]*J[2][2] - J[1][2]*J[2][0])
                                  + J[0][2]*(J[1][0]*J[2][1] - J[1][1]*J[2][0]);
               double vex[3] = {0, 0, 0};
               ic.v0(x, vex);
               double err = 0.0;
               for (int c = 0; c < dim; c++) { err += (v[c] - vex[c])*(v[c] - vex[c]); }
               err = std::sqrt(err);
               const double w = gw[qx]*gw[qy]*((dim == 3) ? gw[qz] : 1.0)*det;
               emax = std::max(emax, err); e1 += w*err; e2 += w*err*err;
            }
   }
   part[3*chunk] = emax; part[3*chunk + 1] = e1; part[3*chunk + 2] = e2;
   });
   out[0] = out[1] = out[2] = 0.0;
   for (int k = 0; k < nchunks; k++) { out[0] = std::max(out[0], part[3*k]); out[1] += part[3*k + 1]; out[2] += part[3*k + 2]; }
}

// `-err` of the reference driver (laghos.cpp:1009-1085): L2 error of the density against the exact Sedov solution at
// time t.  rho = the L2 (Bernstein) density field of ComputeDensity on the current mesh x = S[0 : dim ndofs]; rule
// IntRules.Get(geom, err_order), err_order = 2 max(2 (max(ok, ot) + 1), oq), i.e. err_order/2 + 1 Gauss points per
// axis; at every point (rho_exact(|x_q - 0|) - rho_h(x_q))^2 weighted with w detJ (QuadratureFunction::Integrate).
// Returns the sum of squares of this rank's elements (root after the rank reduction).
inline double sedov_density_error_sum(const Problem &P, const double *S, const double *rho, const SedovExact &sol)
{
   const int dim = P.dim, D = P.D1D, L1 = P.L1D;
   const int err_order = 2*std::max(2*(std::max(P.spec.ok, P.spec.ot) + 1), P.spec.oq);
   const int n = err_order/2 + 1;
   std::vector<double> gx, gw;
   gauss_legendre_01(n, gx, gw);
   std::vector<double> B((size_t)n*D), G((size_t)n*D), BL((size_t)n*L1);
   for (int q = 0; q < n; q++)
   {
      lagrange_eval(P.tab.gll, gx[q], &B[(size_t)q*D], &G[(size_t)q*D]);
      bernstein_eval(L1 - 1, gx[q], &BL[(size_t)q*L1]);
   }
   const int64_t nd = P.ndofs_h1;
   const int nz = (dim == 3) ? n : 1, DZ = (dim == 3) ? D : 1, LZ = (dim == 3) ? L1 : 1;
   std::vector<double> part(256, 0.0);
   int nchunks = 1;
   for_element_chunks(P.NE, nchunks, [&](int chunk, int e_begin, int e_end)
   {
   double sum = 0.0;
   for (int e = e_begin; e < e_end; e++)
   {
      const int *map = &P.h1_map[(size_t)e*P.ND];
      const double *re = rho + (size_t)e*P.NL;
      for (int qz = 0; qz < nz; qz++)
         for (int qy = 0; qy < n; qy++)
            for (int qx = 0; qx < n; qx++)
            {
               double x[3] = {0, 0, 0}, J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
               for (int kz = 0; kz < DZ; kz++)
                  for (int ky = 0; ky < D; ky++)
                     for (int kx = 0; kx < D; kx++)
                     {
                        const double bx = B[(size_t)qx*D + kx], by = B[(size_t)qy*D + ky];
                        const double bz = (dim == 3) ? B[(size_t)qz*D + kz] : 1.0;
                        const double g[3] = {G[(size_t)qx*D + kx]*by*bz, bx*G[(size_t)qy*D + ky]*bz,
                                             (dim == 3) ? bx*by*G[(size_t)qz*D + kz] : 0.0};
                        const int64_t id = map[kx + D*(ky + D*kz)];
                        for (int c = 0; c < dim; c++)
                        {
                           const double xc = S[(size_t)c*nd + id];
                           x[c] += bx*by*bz*xc;
                           for (int d = 0; d < dim; d++) { J[c][d] += g[d]*xc; }
                        }
                     }
               double rh = 0.0;
               for (int lz = 0; lz < LZ; lz++)
                  for (int ly = 0; ly < L1; ly++)
                     for (int lx = 0; lx < L1; lx++)
                     {
                        rh += BL[(size_t)qx*L1 + lx]*BL[(size_t)qy*L1 + ly]*((dim == 3) ? BL[(size_t)qz*L1 + lz] : 1.0)
                              *re[lx + L1*(ly + L1*lz)];
                     }
               const double det = (dim == 2) ? J[0][0]*J[1][1] - J[0][1]*J[1][0]
                                  : J[0][0]*(J[1][1]*J[2][2] - J[1][2]*J[2][1])
                                  - J[0][1]*(J[1][0]*J[2][2] - J[1][2]*J[2][0])
                                  + J[0][2]*(J[1][0]*J[2][1] - J[1][1]*J[2][0]);
               const double r = std::sqrt(x[0]*x[0] + x[1]*x[1] + x[2]*x[2]);
               double rex, vex, pex;
               sol.eval(r, rex, vex, pex);
               const double w = gw[qx]*gw[qy]*((dim == 3) ? gw[qz] : 1.0)*det;
               sum += w*(rex - rh)*(rex - rh);
            }
   }
   part[chunk] = sum;
   });
   double total = 0.0;
   for (int k = 0; k < nchunks; k++) { total += part[k]; }
   return total;
}

} // namespace lagb
