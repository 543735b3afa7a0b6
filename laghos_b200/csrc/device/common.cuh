// Shared device helpers: 1D table struct passed by value (lives in the kernel
// parameter constant bank, so fully unrolled contractions read B/G as constant
// operands), the small tensor contraction, and the per-point QUpdate physics.
#pragma once
#include "qmath.cuh"
#include <stdint.h>

namespace lagb {

// H1: B(q,d), G(q,d) at [q + Q1D*d]; L2 Bernstein: BL(q,l) at [q + Q1D*l]
// (reference laghos_assembly.cpp:153-155, 304-306).
template<int D1D, int Q1D>
struct DevTables
{
   double B[Q1D*D1D];
   double G[Q1D*D1D];
   double BL[Q1D*(D1D > 1 ? D1D - 1 : 1)];
};

// out = T applied along AXIS of in (extents N0,N1,N2, x fastest).
// T is a (Q1D x N1D) table stored [q + Q1D*d].  TR = false: N1D -> Q1D (NA = N1D,
// NB = Q1D); TR = true: Q1D -> N1D (NA = Q1D, NB = N1D).
template<int NA, int NB, int N0, int N1, int N2, int AXIS, bool TR>
__device__ __forceinline__ void contract(const double *T, const double *in, double *out)
{
   constexpr int O0 = (AXIS == 0) ? NB : N0;
   constexpr int O1 = (AXIS == 1) ? NB : N1;
   constexpr int O2 = (AXIS == 2) ? NB : N2;
   for (int k = 0; k < O2; k++)
      for (int j = 0; j < O1; j++)
         for (int i = 0; i < O0; i++)
         {
            double u = 0.0;
            const int b = (AXIS == 0) ? i : (AXIS == 1) ? j : k;
#pragma unroll
            for (int a = 0; a < NA; a++)
            {
               const int s0 = (AXIS == 0) ? a : i;
               const int s1 = (AXIS == 1) ? a : j;
               const int s2 = (AXIS == 2) ? a : k;
               const double t = TR ? T[a + NA*b] : T[b + NB*a];
               u += t*in[s0 + N0*(s1 + N1*s2)];
            }
            out[i + O0*(j + O1*k)] = u;
         }
}

// Programmatic dependent launch (griddepcontrol): a kernel launched with the stream-serialisation attribute may
// be scheduled while its predecessor drains; pdl_wait() blocks until the predecessor grid has completed and its
// writes are visible (a no-op for a normal launch), pdl_launch() lets the successor's launch proceed early.
// Every kernel of the PCG iteration starts with both: ~5 launch gaps of 5-9 us per iteration shrink to the drain.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// running minimum of non-negative doubles (dt estimates): for x, y >= 0 the IEEE bit patterns
// order like unsigned integers, so one atomicMin per CTA replaces a second reduction pass.
// min is exact and order independent: deterministic.
__device__ __forceinline__ void atomic_min_nonneg(double *addr, double v)
{
   atomicMin(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

struct QPointParams
{
   double h0, h1order, inv_h1order, cfl, dt_in;
   int use_viscosity, use_vorticity;
};

// Per-point physics of reference QUpdateBody<DIM> (laghos_solver.cpp:1042-1168).
// J, dV column-major [c + DIM*d]; J0inv as stored in QuadratureData::Jac0inv (read where it is used,
// from shared or global memory: nothing of it is held in registers across the eigen-solve).
// Writes sJ[vd + gd*DIM] = (stress J^-T)(vd,gd) * w * detJ and returns the point's
// dt estimate (already min'ed with p.dt_in).
// Divisions and square roots go through qm::frcp / fdiv / fsqrt (<= 1 ulp, no slow-path calls).
template<int DIM>
__device__ __forceinline__ double qpoint(const double *J, const double *dV, const double e_q,
                                         const double rho0DetJ0w, const double *J0inv,
                                         const double gamma, const double weight, const double inv_weight,
                                         const QPointParams &p, double *sJ)
{
   constexpr int DIM2 = DIM*DIM;
   double Jinv[DIM2], stress[DIM2];
   double detJ;
   if (DIM == 2) { detJ = J[0]*J[3] - J[1]*J[2]; }
   else { detJ = J[0]*(J[4]*J[8] - J[5]*J[7]) + J[3]*(J[2]*J[7] - J[1]*J[8]) + J[6]*(J[1]*J[5] - J[2]*J[4]); }
   const double idetJ = qm::frcp(detJ);
   if (DIM == 2)
   {
      Jinv[0] =  J[3]*idetJ; Jinv[1] = -J[1]*idetJ; Jinv[2] = -J[2]*idetJ; Jinv[3] = J[0]*idetJ;
   }
   else
   {
      Jinv[0] = (J[4]*J[8] - J[5]*J[7])*idetJ;
      Jinv[1] = (J[7]*J[2] - J[8]*J[1])*idetJ;
      Jinv[2] = (J[1]*J[5] - J[2]*J[4])*idetJ;
      Jinv[3] = (J[5]*J[6] - J[3]*J[8])*idetJ;
      Jinv[4] = (J[8]*J[0] - J[6]*J[2])*idetJ;
      Jinv[5] = (J[2]*J[3] - J[0]*J[5])*idetJ;
      Jinv[6] = (J[3]*J[7] - J[4]*J[6])*idetJ;
      Jinv[7] = (J[6]*J[1] - J[7]*J[0])*idetJ;
      Jinv[8] = (J[0]*J[4] - J[1]*J[3])*idetJ;
   }
   const double R = inv_weight*rho0DetJ0w*idetJ;  // reference laghos_solver.cpp:1081 (inv_weight*rho0DetJ0w/detJ)
   const double E = fmax(0.0, e_q);
   const double P = (gamma - 1.0)*R*E;
   const double S = qm::fsqrt(gamma*(gamma - 1.0)*E);
#pragma unroll
   for (int k = 0; k < DIM2; k++) { stress[k] = 0.0; }
#pragma unroll
   for (int d = 0; d < DIM; d++) { stress[d*DIM + d] = -P; }
   double visc_coeff = 0.0;
   if (p.use_viscosity)
   {
      double sg[DIM2];
      // sgrad_v = dV * Jinv   (kernels::Mult summation order: k outer)
#pragma unroll
      for (int j = 0; j < DIM; j++)
#pragma unroll
         for (int i = 0; i < DIM; i++)
         {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; k++) { a += dV[i + k*DIM]*Jinv[k + j*DIM]; }
            sg[i + j*DIM] = a;
         }
      double vorticity_coeff = 1.0;
      if (p.use_vorticity)
      {
         double max_norm = 0.0;
#pragma unroll
         for (int i = 0; i < DIM2; i++) { max_norm = fmax(max_norm, fabs(sg[i])); }
         double grad_norm = 0.0;
         if (max_norm != 0.0)
         {
            double fnorm2 = 0.0;
            const double imax_norm = qm::frcp(max_norm);
#pragma unroll
            for (int i = 0; i < DIM2; i++) { const double en = sg[i]*imax_norm; fnorm2 += en*en; }
            grad_norm = max_norm*qm::fsqrt(fnorm2);
         }
         double tr = 0.0;
#pragma unroll
         for (int i = 0; i < DIM; i++) { tr += sg[i + i*DIM]; }
         const double div_v = fabs(tr);
         vorticity_coeff = (grad_norm > 0.0) ? qm::fdiv(div_v, grad_norm) : 1.0;
      }
      // Symmetrize
#pragma unroll
      for (int i = 0; i < DIM; i++)
#pragma unroll
         for (int j = 0; j < i; j++)
         {
            const double a = 0.5*(sg[i*DIM + j] + sg[j*DIM + i]);
            sg[j*DIM + i] = sg[i*DIM + j] = a;
         }
      double mu, c0, c1, c2 = 0.0;
      if (DIM == 2) { qm::min_eig2(sg[0], sg[2], sg[3], mu, c0, c1); }
      else { qm::min_eig3(sg[0], sg[3], sg[6], sg[4], sg[7], sg[8], mu, c0, c1, c2); }
      // ph_dir = (J * J0inv) * compr_dir, evaluated as J * (J0inv * compr_dir)
      double ph[DIM];
      {
         const double cd[3] = {c0, c1, c2};
         double t[DIM];
#pragma unroll
         for (int i = 0; i < DIM; i++)
         {
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < DIM; j++) { a += J0inv[i + j*DIM]*cd[j]; }
            t[i] = a;
         }
#pragma unroll
         for (int i = 0; i < DIM; i++)
         {
            double a = 0.0;
#pragma unroll
            for (int j = 0; j < DIM; j++) { a += J[i + j*DIM]*t[j]; }
            ph[i] = a;
         }
      }
      const double ph_dir_nl2 = (DIM == 2) ? qm::norml2_2(ph[0], ph[1]) : qm::norml2_3(ph[0], ph[1], ph[DIM-1]);
      const double compr_dir_nl2 = (DIM == 2) ? qm::norml2_2(c0, c1) : qm::norml2_3(c0, c1, c2);
      const double H = p.h0*qm::fdiv(ph_dir_nl2, compr_dir_nl2);
      visc_coeff = 2.0*R*H*H*fabs(mu);
      const double eps = 1e-12;
      visc_coeff += 0.5*R*H*S*vorticity_coeff*(1.0 - qm::smooth_step_01(mu - 2.0*eps, eps));
#pragma unroll
      for (int k = 0; k < DIM2; k++) { stress[k] = stress[k] + visc_coeff*sg[k]; }
   }
   const double sv = (DIM == 2) ? qm::min_sv2(J[0], J[1], J[2], J[3])
                     : qm::min_sv3(J[0], J[1], J[2], J[3], J[DIM2 > 4 ? 4 : 0], J[DIM2 > 5 ? 5 : 0],
                                   J[DIM2 > 6 ? 6 : 0], J[DIM2 > 7 ? 7 : 0], J[DIM2 > 8 ? 8 : 0]);
   const double h_min = sv*p.inv_h1order;
   const double ih_min = qm::frcp(h_min);
   const double irho_ih_min_sq = qm::fdiv(ih_min*ih_min, R);
   const double idt = S*ih_min + 2.5*visc_coeff*irho_ih_min_sq;
   double dt_q = p.dt_in;
   if (detJ < 0.0) { dt_q = 0.0; }
   else if (idt > 0.0) { dt_q = fmin(dt_q, qm::fdiv(p.cfl, idt)); }
   // stressJiT = stress * Jinv^T, scaled
   const double wd = weight*detJ;
#pragma unroll
   for (int j = 0; j < DIM; j++)
#pragma unroll
      for (int i = 0; i < DIM; i++)
      {
         double a = 0.0;
#pragma unroll
         for (int k = 0; k < DIM; k++) { a += stress[i + k*DIM]*Jinv[j + k*DIM]; }
         sJ[i + j*DIM] = a*wd;
      }
   return dt_q;
}

} // namespace lagb
