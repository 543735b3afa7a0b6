// Tuned 3D velocity-mass partial-assembly apply:  y += G^t B^t D B G x
// (reference MassPAOperator::Mult -> MFEM MassIntegrator::AddMultPA,
//  laghos_assembly.cpp:117-121; arithmetic as in amr/laghos_assembly.cpp:878-963).
//
// B200 mapping (DESIGN.md "mass3d"):
//   * NC*D1D threads per element, NB elements per CTA, NC velocity components per pass
//     (NC = 3 in the batched PCG: the quadrature data D is read ONCE for all three
//     component solves of SolveVelocity, laghos_solver.cpp:363-398).
//   * the D values of the thread's phase-B columns are requested first (registers) so that
//     their DRAM latency overlaps the gather and phase A;
//   * phase 0  the CTA's restriction indices are staged in shared memory (and, in the staged
//              variant DIRECT_GATHER = false, the dof values too, lanes along the dof index);
//   * phase A  thread (c,e,dz) gathers slice dz of component c, contracts x then y in registers
//              (B in the kernel-parameter constant bank, fully unrolled) and stores the
//              Q1D^2 plane to shared memory;
//   * phase B  flat (element,column) index over the CTA's NB*Q1D^2 quadrature columns:
//              contract z (D1D -> Q1D), scale by D, contract z back (Q1D -> D1D), in registers,
//              in place in shared memory; optionally accumulates d^t A d = sum_q D u_q^2 for the
//              PCG denominator at no extra memory traffic;
//   * phase C  thread (c,e,dz) contracts y then x back to dofs and scatter-adds them
//              (red.global.add.f64) straight from registers (DIRECT_SCATTER) or through a
//              cooperative phase D with lanes along the dof index.
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
   static constexpr int TG = NB*D1D;                // threads per component group (element, slice)
   static constexpr int TA = NC*TG;                 // working threads (component, element, slice)
   static constexpr int T = ((TA + 31)/32)*32;      // CTA size, padded to whole warps
   static constexpr int NCOL = (NB*QQ + T - 1)/T;   // quadrature columns per thread in phase B
   static constexpr bool PREFETCH = (NCOL*Q1D <= 24); // hold the columns' D values in registers across phase A
   static constexpr int PLANE = ((QQ > DD ? QQ : DD) | 1);   // plane stride (odd: conflict-free 64-bit)
   static constexpr int SMEM_DOUBLES = NC*NB*D1D*PLANE;
   // index stride per slice: odd (conflict-free scalar reads, lanes = slices) for the batched
   // apply; dense for one component, where 128-bit index loads measured faster (187 vs 210 us)
   static constexpr int IDXS = (NC == 1) ? DD : (DD | 1);
   static constexpr size_t SMEM_BYTES = (size_t)SMEM_DOUBLES*sizeof(double) + (size_t)NB*D1D*IDXS*sizeof(int);
};

template<int D1D, int Q1D, int NB, int NC, bool WITH_DEN, int MINB, bool DIRECT_SCATTER, bool DIRECT_GATHER = false>
__global__ void __launch_bounds__(((NC*NB*D1D + 31)/32)*32, MINB)
mass3d(const __grid_constant__ DevTables<D1D,Q1D> tab, const int NE, const int64_t cstride,
       const int *__restrict__ map, const double *__restrict__ Dq,
       const double *__restrict__ x, double *__restrict__ y, double *__restrict__ den_part)
{
   using C = Mass3DCfg<D1D,Q1D,NB,NC>;
   pdl_launch();                    // the PCG's next kernel may be staged while this grid drains
   extern __shared__ double sV[];   // [c][e_loc][dz][PLANE]
   int *sIdx = reinterpret_cast<int*>(sV + C::SMEM_DOUBLES);   // [e_loc][dz][IDXS]
   const int t = threadIdx.x;
   const int c = t / C::TG, r = t - c*C::TG;
   const int e_loc = r / D1D, dz = r % D1D;
   const int eb = blockIdx.x*NB;
   const int nel = min(NB, NE - eb);
   const bool active = (t < C::TA) && (e_loc < nel);
   const int ncols = nel*C::QQ;

   // quadrature data of this thread's phase-B columns: issued first so that the DRAM
   // latency overlaps the gather and phase A
   double dq[C::PREFETCH ? C::NCOL : 1][Q1D];
#pragma unroll
   for (int k = 0; k < (C::PREFETCH ? C::NCOL : 0); k++)
   {
      const int f = t + k*C::T;
      if (f < ncols)
      {
         const int e2 = f / C::QQ, col = f - e2*C::QQ;
         const double *dptr = Dq + (size_t)(eb + e2)*C::NQ + col;
#pragma unroll
         for (int qz = 0; qz < Q1D; qz++) { dq[k][qz] = __ldg(dptr + C::QQ*qz); }
      }
   }
   pdl_wait();                      // x (and the zero-filled y) come from the predecessor kernel; D and the map do not
   // ---- phase 0: cooperative gather, lanes along the element-local dof index (runs of D1D
   //      contiguous L-vector entries per lattice row).  Slice (c,e,dz) is parked in the
   //      first DD slots of its own plane.
   if (DIRECT_GATHER)
   {
      // only the restriction indices are staged; each slice thread gathers its own DD values
      for (int it = t; it < nel*C::ND; it += C::T) { sIdx[(it / C::DD)*C::IDXS + it % C::DD] = __ldg(map + (size_t)eb*C::ND + it); }
   }
   else
   {
      // all index loads first, then all value loads (independent requests in flight), then the stores
      constexpr int NIT = (NB*C::ND + C::T - 1)/C::T;
      const int nd = nel*C::ND;
      int id[NIT];
#pragma unroll
      for (int k = 0; k < NIT; k++)
      {
         const int it = t + k*C::T;
         id[k] = (it < nd) ? __ldg(map + (size_t)eb*C::ND + it) : 0;
      }
      double xv[NIT][NC];
#pragma unroll
      for (int k = 0; k < NIT; k++)
      {
         const int it = t + k*C::T;
#pragma unroll
         for (int cc = 0; cc < NC; cc++) { xv[k][cc] = (it < nd) ? x[(size_t)cc*cstride + id[k]] : 0.0; }
      }
#pragma unroll
      for (int k = 0; k < NIT; k++)
      {
         const int it = t + k*C::T;
         if (it < nd)
         {
            const int e2 = it / C::ND, i = it - e2*C::ND;
            const int z = i / C::DD, ixy = i - z*C::DD;
            sIdx[(it / C::DD)*C::IDXS + it % C::DD] = id[k];
#pragma unroll
            for (int cc = 0; cc < NC; cc++) { sV[((size_t)(cc*NB + e2)*D1D + z)*C::PLANE + ixy] = xv[k][cc]; }
         }
      }
   }
   __syncthreads();
   double *pl = sV + ((size_t)(c*NB + e_loc)*D1D + dz)*C::PLANE;
   // ---- phase A: x then y contraction of the slice, store plane ----
   if (active)
   {
      double XG[DIRECT_GATHER ? C::DD : 1];
      if (DIRECT_GATHER)
      {
         const int *ids = sIdx + (e_loc*D1D + dz)*C::IDXS;
         const double *xc = x + (size_t)c*cstride;
#pragma unroll
         for (int i = 0; i < C::DD; i++) { XG[i] = xc[ids[i]]; }
      }
      double U[Q1D][D1D];
#pragma unroll
      for (int dy = 0; dy < D1D; dy++)
      {
         double X[D1D];
#pragma unroll
         for (int dx = 0; dx < D1D; dx++) { X[dx] = DIRECT_GATHER ? XG[dx + D1D*dy] : pl[dx + D1D*dy]; }
#pragma unroll
         for (int qx = 0; qx < Q1D; qx++)
         {
            double u = 0.0;
#pragma unroll
            for (int dx = 0; dx < D1D; dx++) { u += tab.B[qx + Q1D*dx]*X[dx]; }
            U[qx][dy] = u;
         }
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
   double den[NC];
#pragma unroll
   for (int cc = 0; cc < NC; cc++) { den[cc] = 0.0; }
#pragma unroll
   for (int k = 0; k < C::NCOL; k++)
   {
      const int f = t + k*C::T;
      if (f < ncols)
      {
         const int e2 = f / C::QQ, col = f - e2*C::QQ;
         const int kd = C::PREFETCH ? k : 0;
         if (!C::PREFETCH)
         {
            const double *dptr = Dq + (size_t)(eb + e2)*C::NQ + col;
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++) { dq[0][qz] = __ldg(dptr + C::QQ*qz); }
         }
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
               const double dw = dq[kd][qz]*w;
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
   __syncthreads();
   // ---- phase C: y then x back; the slice result overwrites the first DD plane slots ----
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
      if (DIRECT_SCATTER)
      {
         // scatter-add straight from registers (restriction indices of the slice from smem)
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
      else
      {
         // all plane reads of this thread are complete (Z holds them): safe to overwrite
#pragma unroll
         for (int dy = 0; dy < D1D; dy++)
#pragma unroll
            for (int dx = 0; dx < D1D; dx++)
            {
               double o = 0.0;
#pragma unroll
               for (int qx = 0; qx < Q1D; qx++) { o += tab.B[qx + Q1D*dx]*Z[qx][dy]; }
               pl[dx + D1D*dy] = o;
            }
      }
   }
   if (!DIRECT_SCATTER)
   {
      __syncthreads();
      // ---- phase D: cooperative scatter-add (red.global.add.f64), lanes along the dof index ----
      for (int it = t; it < nel*C::ND; it += C::T)
      {
         const int e2 = it / C::ND, i = it - e2*C::ND;
         const int id = sIdx[(it / C::DD)*C::IDXS + it % C::DD];
         const int z = i / C::DD, ixy = i - z*C::DD;
#pragma unroll
         for (int cc = 0; cc < NC; cc++)
         {
            atomicAdd(y + (size_t)cc*cstride + id, sV[((size_t)(cc*NB + e2)*D1D + z)*C::PLANE + ixy]);
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
