// Tuned 3D velocity-mass partial-assembly apply:  y += G^t B^t D B G x
// (reference MassPAOperator::Mult -> MFEM MassIntegrator::AddMultPA,
//  laghos_assembly.cpp:117-121; arithmetic as in amr/laghos_assembly.cpp:878-963).
//
// B200 mapping (DESIGN.md "mass3d"):
//   * D1D threads per element, NB elements per CTA, NC velocity components per pass
//     (NC = 3 in the batched PCG: the quadrature data D is read ONCE for all three
//     component solves of SolveVelocity, laghos_solver.cpp:363-398).
//   * phase A  thread (e,dz) gathers the xy-slice dz of the element straight from the
//              L-vector through the restriction map, contracts x then y in registers
//              (B in the kernel-parameter constant bank, fully unrolled) and stores the
//              Q1D^2 plane to shared memory;
//   * phase B  flat (element,column) index over the CTA's NB*Q1D^2 quadrature columns:
//              contract z (D1D -> Q1D), scale by D (coalesced streaming loads, each byte
//              of D touched once), contract z back (Q1D -> D1D), in registers, in place
//              in shared memory; optionally accumulates d^t A d = sum_q D u_q^2 for the
//              PCG denominator at no extra memory traffic;
//   * phase C  thread (e,dz) contracts y then x back to dofs and scatter-adds to the
//              L-vector (red.global.add.f64).
//   Shared memory traffic is 4*D1D*Q1D^2 doubles per element and component (no
//   per-FMA shared operands); every thread is active in every phase.
#pragma once
#include "common.cuh"

namespace lagb {
namespace tuned {

template<int D1D, int Q1D, int NB, int NC>
struct Mass3DCfg
{
   static constexpr int DD = D1D*D1D, QQ = Q1D*Q1D, ND = D1D*DD, NQ = Q1D*QQ;
   static constexpr int TA = NB*D1D;                // working threads (element, slice)
   static constexpr int T = ((TA + 31)/32)*32;      // CTA size, padded to whole warps
   static constexpr int PLANE = QQ + 1;             // padded plane stride (odd: conflict-free 64-bit)
   static constexpr int SMEM_DOUBLES = NC*NB*D1D*PLANE;
   static constexpr size_t SMEM_BYTES = (size_t)SMEM_DOUBLES*sizeof(double);
};

template<int D1D, int Q1D, int NB, int NC, bool WITH_DEN>
__global__ void __launch_bounds__(((NB*D1D + 31)/32)*32)
mass3d(const __grid_constant__ DevTables<D1D,Q1D> tab, const int NE, const int64_t cstride,
       const int *__restrict__ map, const double *__restrict__ Dq,
       const double *__restrict__ x, double *__restrict__ y, double *__restrict__ den_part)
{
   using C = Mass3DCfg<D1D,Q1D,NB,NC>;
   extern __shared__ double sV[];   // [c][e_loc][dz][PLANE]
   const int t = threadIdx.x;
   const int e_loc = t / D1D, dz = t % D1D;
   const int eb = blockIdx.x*NB;
   const int e = eb + e_loc;
   const bool active = (t < C::TA) && (e < NE);

   int idx[C::DD];
   if (active)
   {
      const int *m = map + (size_t)e*C::ND + dz*C::DD;
#pragma unroll
      for (int i = 0; i < C::DD; i++) { idx[i] = m[i]; }
   }
   // ---- phase A: gather slice, x then y contraction, store plane ----
#pragma unroll
   for (int c = 0; c < NC; c++)
   {
      double U[Q1D][D1D];
      if (active)
      {
         const double *xc = x + (size_t)c*cstride;
#pragma unroll
         for (int dy = 0; dy < D1D; dy++)
         {
            double X[D1D];
#pragma unroll
            for (int dx = 0; dx < D1D; dx++) { X[dx] = xc[idx[dx + D1D*dy]]; }
#pragma unroll
            for (int qx = 0; qx < Q1D; qx++)
            {
               double u = 0.0;
#pragma unroll
               for (int dx = 0; dx < D1D; dx++) { u += tab.B[qx + Q1D*dx]*X[dx]; }
               U[qx][dy] = u;
            }
         }
      }
      double *pl = sV + ((size_t)(c*NB + e_loc)*D1D + dz)*C::PLANE;
      if (active)
      {
#pragma unroll
         for (int qx = 0; qx < Q1D; qx++)
#pragma unroll
            for (int qy = 0; qy < Q1D; qy++)
            {
               double v = 0.0;
#pragma unroll
               for (int dy = 0; dy < D1D; dy++) { v += tab.B[qy + Q1D*dy]*U[qx][dy]; }
               pl[qx + Q1D*qy] = v;
            }
      }
   }
   __syncthreads();
   // ---- phase B: z contraction, scale by D, z back (flat element-column index) ----
   double den[NC];
#pragma unroll
   for (int c = 0; c < NC; c++) { den[c] = 0.0; }
   {
      const int ncols = min(NB, NE - eb)*C::QQ;
      for (int f = t; f < ncols; f += C::T)
      {
         const int e2 = f / C::QQ, col = f - e2*C::QQ;
         const double *dptr = Dq + (size_t)(eb + e2)*C::NQ + col;
         double dq[Q1D];
#pragma unroll
         for (int qz = 0; qz < Q1D; qz++) { dq[qz] = __ldg(dptr + C::QQ*qz); }
#pragma unroll
         for (int c = 0; c < NC; c++)
         {
            double *colp = sV + ((size_t)(c*NB + e2)*D1D)*C::PLANE + col;
            double V[D1D], W[Q1D];
#pragma unroll
            for (int k = 0; k < D1D; k++) { V[k] = colp[k*C::PLANE]; }
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++)
            {
               double w = 0.0;
#pragma unroll
               for (int k = 0; k < D1D; k++) { w += tab.B[qz + Q1D*k]*V[k]; }
               const double dw = dq[qz]*w;
               if (WITH_DEN) { den[c] += dw*w; }
               W[qz] = dw;
            }
#pragma unroll
            for (int k = 0; k < D1D; k++)
            {
               double v = 0.0;
#pragma unroll
               for (int qz = 0; qz < Q1D; qz++) { v += tab.B[qz + Q1D*k]*W[qz]; }
               colp[k*C::PLANE] = v;
            }
         }
      }
   }
   __syncthreads();
   // ---- phase C: y then x back, scatter-add ----
#pragma unroll
   for (int c = 0; c < NC; c++)
   {
      if (active)
      {
         const double *pl = sV + ((size_t)(c*NB + e_loc)*D1D + dz)*C::PLANE;
         double Z[Q1D][D1D];
#pragma unroll
         for (int qx = 0; qx < Q1D; qx++)
         {
            double P[Q1D];
#pragma unroll
            for (int qy = 0; qy < Q1D; qy++) { P[qy] = pl[qx + Q1D*qy]; }
#pragma unroll
            for (int dy = 0; dy < D1D; dy++)
            {
               double z = 0.0;
#pragma unroll
               for (int qy = 0; qy < Q1D; qy++) { z += tab.B[qy + Q1D*dy]*P[qy]; }
               Z[qx][dy] = z;
            }
         }
         double *yc = y + (size_t)c*cstride;
#pragma unroll
         for (int dy = 0; dy < D1D; dy++)
#pragma unroll
            for (int dx = 0; dx < D1D; dx++)
            {
               double o = 0.0;
#pragma unroll
               for (int qx = 0; qx < Q1D; qx++) { o += tab.B[qx + Q1D*dx]*Z[qx][dy]; }
               atomicAdd(yc + idx[dx + D1D*dy], o);
            }
      }
   }
   if (WITH_DEN)
   {
      // deterministic block reduction: warp shuffle tree, then warp partials in order
      __syncthreads();
      double *red = sV;
      constexpr int NW = (C::T + 31)/32;
#pragma unroll
      for (int c = 0; c < NC; c++)
      {
         double v = den[c];
         for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
         if ((t & 31) == 0) { red[c*NW + (t >> 5)] = v; }
      }
      __syncthreads();
      if (t < NC)
      {
         double s = 0.0;
         for (int w = 0; w < NW; w++) { s += red[t*NW + w]; }
         den_part[(size_t)blockIdx.x*NC + t] = s;
      }
   }
}

} // namespace tuned
} // namespace lagb
