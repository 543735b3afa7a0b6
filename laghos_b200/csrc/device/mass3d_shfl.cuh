// EXPERIMENTAL (selected only with lagb_tune_set(ctx, 0, 5)).  Measured on B200, cube01_hex -rs 5 Q3Q2, NC = 3:
// correct (checksum identical to the production kernel) but SLOWER: 488 us vs 339 us.  The whole plane lives in
// registers (168 per thread -> 4 CTAs = 12 warps per SM, against 18 warps at 96 registers) and the
// butterflies add 72 selects + 72 SHFL per transpose; the shared-memory planes stay the production path.
//
// mass3d with a register/shuffle hand-off between the slice phases (A, C) and the column phase (B)
// instead of shared-memory planes.  Motivation (profiles/ncu_mass3d_r1_final.txt): the LSU data pipe
// is the binding resource of mass3d (83 %); the plane stores / loads are ~45 % of its wavefronts.
//
// D1D = 4 only: the 4 slice threads (dz = 0..3) of one (component, element) are 4 adjacent lanes.
// Each lane holds its plane V[col], col = qx + Q1D*qy, in registers.  Columns are dealt to the 4 lanes
// round-robin (block m = {4k + m}); a two-step xor butterfly (lane^1, lane^2) with compile-time
// register indices and selects transposes the 4 x 4 block matrix, after which lane j holds, for its
// columns 4k + j, the values of all four slices (slot m = slice m).  Phase B runs in registers, the
// same butterfly (an involution) brings the planes back.  Per transpose: 36 64-bit shuffles and
// 72 selects per thread against 36 STS.64 + 36 LDS.64.
#pragma once
#include "common.cuh"

namespace lagb {
namespace tuned {

// in-register transpose of the 4 x 4 block matrix held by 4 adjacent lanes (block m = V[4k + m])
template<int NK>
__device__ __forceinline__ void quad_transpose(double (&V)[4*NK], const int dz)
{
   const bool b0 = (dz & 1) != 0, b1 = (dz & 2) != 0;
#pragma unroll
   for (int p = 0; p < 4; p += 2)          // step 1: partner lane^1 exchanges block p+1 (b0 = 0) / p (b0 = 1)
#pragma unroll
      for (int k = 0; k < NK; k++)
      {
         const double send = b0 ? V[4*k + p] : V[4*k + p + 1];
         const double recv = __shfl_xor_sync(0xffffffffu, send, 1);
         if (b0) { V[4*k + p] = recv; } else { V[4*k + p + 1] = recv; }
      }
#pragma unroll
   for (int s = 0; s < 2; s++)             // step 2: partner lane^2 exchanges slots {2,3} (b1 = 0) / {0,1} (b1 = 1)
#pragma unroll
      for (int k = 0; k < NK; k++)
      {
         const double send = b1 ? V[4*k + s] : V[4*k + 2 + s];
         const double recv = __shfl_xor_sync(0xffffffffu, send, 2);
         if (b1) { V[4*k + s] = recv; } else { V[4*k + 2 + s] = recv; }
      }
}

template<int Q1D, int NB, int NC, bool WITH_DEN, int MINB>
__global__ void __launch_bounds__(NC*NB*4, MINB)
mass3d_shfl(const __grid_constant__ DevTables<4,Q1D> tab, const int NE, const int64_t cstride,
            const int *__restrict__ map, const double *__restrict__ Dq,
            const double *__restrict__ x, double *__restrict__ y, double *__restrict__ den_part)
{
   constexpr int D1D = 4, DD = 16, ND = 64, QQ = Q1D*Q1D, NQ = QQ*Q1D, NK = QQ/4, T = NC*NB*4;
   static_assert(QQ % 4 == 0 && T % 32 == 0, "quad hand-off needs Q1D^2 % 4 == 0 and whole warps");
   __shared__ double red[NC*(T/32)];
   const int t = threadIdx.x;
   const int c = t / (NB*4), r = t - c*(NB*4);
   const int e_loc = r >> 2, dz = r & 3;
   const int eb = blockIdx.x*NB;
   const bool active = (eb + e_loc) < NE;
   const int e = active ? eb + e_loc : NE - 1;        // inactive lanes compute on a valid element, store nothing
   // restriction indices of the slice (needed again for the scatter)
   int idx[DD];
   {
      const int *m = map + (size_t)e*ND + dz*DD;
#pragma unroll
      for (int i = 0; i < DD; i++) { idx[i] = __ldg(m + i); }
   }
   // ---- phase A: gather, x then y contraction, plane in registers ----
   double V[QQ];
   {
      const double *xc = x + (size_t)c*cstride;
      double U[Q1D][D1D];
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
#pragma unroll
      for (int qx = 0; qx < Q1D; qx++)
#pragma unroll
         for (int qy = 0; qy < Q1D; qy++)
         {
            double v = 0.0;
#pragma unroll
            for (int dy = 0; dy < D1D; dy++) { v += tab.B[qy + Q1D*dy]*U[qx][dy]; }
            V[qx + Q1D*qy] = v;
         }
   }
   quad_transpose<NK>(V, dz);     // lane dz now owns columns 4k + dz; V[4k + m] = slice m of that column
   // ---- phase B: z contraction, scale by D, z back, per owned column ----
   double den = 0.0;
   {
      const double *dcol = Dq + (size_t)e*NQ + dz;
#pragma unroll
      for (int k = 0; k < NK; k++)
      {
         double dq[Q1D], W[Q1D];
#pragma unroll
         for (int qz = 0; qz < Q1D; qz++) { dq[qz] = __ldg(dcol + 4*k + QQ*qz); }
#pragma unroll
         for (int qz = 0; qz < Q1D; qz++)
         {
            double w = 0.0;
#pragma unroll
            for (int m = 0; m < D1D; m++) { w += tab.B[qz + Q1D*m]*V[4*k + m]; }
            const double dw = dq[qz]*w;
            if (WITH_DEN) { den += dw*w; }
            W[qz] = dw;
         }
#pragma unroll
         for (int m = 0; m < D1D; m++)
         {
            double v = 0.0;
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++) { v += tab.B[qz + Q1D*m]*W[qz]; }
            V[4*k + m] = v;
         }
      }
   }
   quad_transpose<NK>(V, dz);     // back: V[col] is the plane of slice dz again
   // ---- phase C: y then x back, scatter-add from registers ----
   {
      double Z[Q1D][D1D];
#pragma unroll
      for (int qx = 0; qx < Q1D; qx++)
#pragma unroll
         for (int dy = 0; dy < D1D; dy++)
         {
            double z = 0.0;
#pragma unroll
            for (int qy = 0; qy < Q1D; qy++) { z += tab.B[qy + Q1D*dy]*V[qx + Q1D*qy]; }
            Z[qx][dy] = z;
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
            if (active) { atomicAdd(yc + idx[dx + D1D*dy], o); }
         }
   }
   if (WITH_DEN)
   {
      // the den of a column is owned by exactly one lane; inactive lanes contribute nothing
      double v = active ? den : 0.0;
      for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
      constexpr int NWC = (NB*4)/32 > 0 ? (NB*4)/32 : 1;     // warps per component group (NB*4 >= 32)
      if ((t & 31) == 0) { red[t >> 5] = v; }
      __syncthreads();
      if (t < NC)
      {
         double s = 0.0;
         for (int w = 0; w < NWC; w++) { s += red[t*NWC + w]; }
         den_part[(size_t)blockIdx.x*NC + t] = s;
      }
   }
}

} // namespace tuned
} // namespace lagb
