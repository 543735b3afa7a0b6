// Persistent form of the H1 mass apply of device/mass3d.cuh (direct gather / direct scatter layout, same phases,
// same arithmetic): a CTA walks over element batches b = blockIdx.x, + gridDim.x, ... instead of one batch per CTA.
//
// ncu of the one-batch-per-CTA kernel (profiles/ncu_mass3d_r2.txt, stall sampling): 10.6 % of the samples wait for
// the restriction indices at the very start of every 9-us CTA (global load -> shared store), before anything else
// can be issued, and the quadrature data of the column phase is requested at the same moment.  Here
//   * the indices of batch b+1 stream into the second half of a double-buffered index array with 4-byte cp.async
//     (LDGSTS) while batch b is processed: no register cost;
//   * the D values of batch b+1's columns are requested right after phase B of batch b has consumed the current
//     ones (same registers);
//   * tables / parameters are set up once, and the d^t A d partials are accumulated over the CTA's batches
//     (grid * NC partials instead of NE/NB * NC).
// Everything else (slice threads, plane layout, in-register contractions, red.global.add.f64 scatter) is mass3d's.
#pragma once
#include "mass3d.cuh"
#include "staged3d.cuh"

namespace lagb {
namespace tuned {

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
   const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
   asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(gsrc) : "memory");
}

template<int D1D, int Q1D, int NB, int NC, bool WITH_DEN, int MINB>
__global__ void __launch_bounds__(((NC*NB*D1D + 31)/32)*32, MINB)
mass3d_persist(const __grid_constant__ DevTables<D1D,Q1D> tab, const int NE, const int64_t cstride,
               const int *__restrict__ map, const double *__restrict__ Dq,
               const double *__restrict__ x, double *__restrict__ y, double *__restrict__ den_part)
{
   using C = Mass3DCfg<D1D,Q1D,NB,NC>;
   static_assert(C::PREFETCH, "column D values are held in registers");
   pdl_launch();
   extern __shared__ double sV[];   // [c][e_loc][dz][PLANE]
   constexpr int IDXN = NB*D1D*C::IDXS;                      // ints per index buffer
   int *sIdx0 = reinterpret_cast<int*>(sV + C::SMEM_DOUBLES); // [2][e_loc][dz][IDXS]
   const int t = threadIdx.x;
   const int c = t / C::TG, r = t - c*C::TG;
   const int e_loc = r / D1D, dz = r % D1D;
   const int nbatch = (NE + NB - 1)/NB;
   double *pl = sV + ((size_t)(c*NB + e_loc)*D1D + dz)*C::PLANE;
   double den[NC];
#pragma unroll
   for (int cc = 0; cc < NC; cc++) { den[cc] = 0.0; }

   auto load_dq = [&](int b, double (&dq)[C::NCOL][Q1D])
   {
      const int eb = b*NB, ncols = min(NB, NE - eb)*C::QQ;
#pragma unroll
      for (int k = 0; k < C::NCOL; k++)
      {
         const int f = t + k*C::T;
         if (b < nbatch && f < ncols)
         {
            const int e2 = f / C::QQ, col = f - e2*C::QQ;
            const double *dptr = Dq + (size_t)(eb + e2)*C::NQ + col;
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++) { dq[k][qz] = __ldg(dptr + C::QQ*qz); }
         }
      }
   };
   auto stage_idx = [&](int b, int *dst)                      // asynchronous: batch b's restriction indices -> dst
   {
      if (b < nbatch)
      {
         const int eb = b*NB, nd = min(NB, NE - eb)*C::ND;
         for (int it = t; it < nd; it += C::T) { cp_async4(dst + (it / C::DD)*C::IDXS + it % C::DD, map + (size_t)eb*C::ND + it); }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
   };

   double dq[C::NCOL][Q1D];
   int b = blockIdx.x;
   load_dq(b, dq);                  // D and the map do not depend on the predecessor kernel
   stage_idx(b, sIdx0);
   pdl_wait();                      // x (and the zero-filled y) do
   int cur = 0;
   for (; b < nbatch; b += gridDim.x, cur ^= 1)
   {
      const int eb = b*NB;
      const int nel = min(NB, NE - eb);
      const bool active = (t < C::TA) && (e_loc < nel);
      const int ncols = nel*C::QQ;
      int *sIdx = sIdx0 + cur*IDXN;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();             // this batch's indices are in place; everyone is done with the previous batch
      stage_idx(b + gridDim.x, sIdx0 + (cur ^ 1)*IDXN);
      // ---- phase A: gather the slice, x then y contraction, store plane ----
      if (active)
      {
         double XG[C::DD];
         const int *ids = sIdx + (e_loc*D1D + dz)*C::IDXS;
         const double *xc = x + (size_t)c*cstride;
#pragma unroll
         for (int i = 0; i < C::DD; i++) { XG[i] = xc[ids[i]]; }
         double U[Q1D][D1D];
#pragma unroll
         for (int dy = 0; dy < D1D; dy++)
#pragma unroll
            for (int qx = 0; qx < Q1D; qx++)
            {
               double u = 0.0;
#pragma unroll
               for (int dx = 0; dx < D1D; dx++) { u += tab.B[qx + Q1D*dx]*XG[dx + D1D*dy]; }
               U[qx][dy] = u;
            }
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
      __syncthreads();
      // ---- phase B: z contraction, scale by D, z back (flat element-column index) ----
#pragma unroll
      for (int k = 0; k < C::NCOL; k++)
      {
         const int f = t + k*C::T;
         if (f < ncols)
         {
            const int e2 = f / C::QQ, col = f - e2*C::QQ;
#pragma unroll
            for (int cc = 0; cc < NC; cc++)
            {
               double *colp = sV + ((size_t)(cc*NB + e2)*D1D)*C::PLANE + col;
               double V[D1D], W[Q1D];
#pragma unroll
               for (int kk = 0; kk < D1D; kk++) { V[kk] = colp[kk*C::PLANE]; }
#pragma unroll
               for (int qz = 0; qz < Q1D; qz++)
               {
                  double w = 0.0;
#pragma unroll
                  for (int kk = 0; kk < D1D; kk++) { w += tab.B[qz + Q1D*kk]*V[kk]; }
                  const double dw = dq[k][qz]*w;
                  if (WITH_DEN) { den[cc] += dw*w; }
                  W[qz] = dw;
               }
#pragma unroll
               for (int kk = 0; kk < D1D; kk++)
               {
                  double v = 0.0;
#pragma unroll
                  for (int qz = 0; qz < Q1D; qz++) { v += tab.B[qz + Q1D*kk]*W[qz]; }
                  colp[kk*C::PLANE] = v;
               }
            }
         }
      }
      load_dq(b + gridDim.x, dq);   // next batch's quadrature data: in flight during phase C and the next phase A
      __syncthreads();
      // ---- phase C: y then x back, scatter-add straight from registers ----
      if (active)
      {
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
         const int *ids = sIdx + (e_loc*D1D + dz)*C::IDXS;
         double *yc = y + (size_t)c*cstride;
#pragma unroll
         for (int dy = 0; dy < D1D; dy++)
#pragma unroll
            for (int dx = 0; dx < D1D; dx++)
            {
               double o = 0.0;
#pragma unroll
               for (int qx = 0; qx < Q1D; qx++) { o += tab.B[qx + Q1D*dx]*Z[qx][dy]; }
               atomicAdd(yc + ids[dx + D1D*dy], o);
            }
      }
   }
   asm volatile("cp.async.wait_group 0;" ::: "memory");
   if (WITH_DEN)
   {
      __syncthreads();
      double *red = sV;
      constexpr int NW = (C::T + 31)/32;
#pragma unroll
      for (int cc = 0; cc < NC; cc++)
      {
         double v = den[cc];
         for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
         if ((t & 31) == 0) { red[cc*NW + (t >> 5)] = v; }
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
