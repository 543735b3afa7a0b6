// H1 mass partial-assembly apply for the high orders (D1D >= 5):  y += G^t B^t D B G x
// (reference MassPAOperator::Mult -> MassIntegrator::AddMultPA, laghos_assembly.cpp:117-121).
//
// The slice kernel of device/mass3d.cuh keeps a Q1D x D1D tile per thread in registers and has only
// NC*D1D threads per element: at Q4Q3 / Q5Q4 that is 40-60 doubles per tile and 12-14 resident warps per SM
// (measured 22 % / 12 % of the HBM peak).  Here every 1D contraction is a flat list of pencils over all
// threads of the CTA, as in the Force / L2-mass kernels of staged3d.cuh: a pencil holds D1D + Q1D values,
// shared memory sees (D1D + Q1D) accesses per D1D*Q1D FMAs (0.27 per FMA at Q5Q4), and the quadrature
// coefficient of a column is loaded once for the NC components (z pencils loop over the components).
// Rows are padded to odd strides (DP, QP) so that lanes along the row index hit distinct banks.
#pragma once
#include "staged3d.cuh"
#include "mass3d.cuh"

namespace lagb {
namespace tuned {

template<int D1D, int Q1D, int NC>
struct MassPencilCfg
{
   static constexpr int DD = D1D*D1D, QQ = Q1D*Q1D, ND = D1D*DD, NQ = Q1D*QQ;
   static constexpr int QP = Q1D | 1, DP = D1D | 1;
   static constexpr int S_X = DD*DP;            // dofs [dz][dy][dx], rows padded
   static constexpr int S_T1 = DD*QP;           // [dz][dy][qx], rows padded
   static constexpr int S_T2 = D1D*QQ;          // [dz][qy][qx]
   static constexpr int PER_EC = S_X + S_T1 + S_T2;   // per (element, component)
};

template<int D1D, int Q1D, int NB, int NC, int NT, bool WITH_DEN>
__global__ void __launch_bounds__(NT)
mass3d_pencil(const __grid_constant__ DevTables<D1D,Q1D> tab, const int NE, const int64_t cstride,
              const int *__restrict__ map, const double *__restrict__ Dq,
              const double *__restrict__ x, double *__restrict__ y, double *__restrict__ den_part)
{
   using C = MassPencilCfg<D1D,Q1D,NC>;
   pdl_launch();
   extern __shared__ double smem[];
   constexpr int PE = C::PER_EC, DD = C::DD, QQ = C::QQ, QP = C::QP, DP = C::DP;
   const int tid = threadIdx.x;
   const int eb = blockIdx.x*NB;
   const int nel = min(NB, NE - eb);
   // element-component slot ec = e*NC + c: X | T1 | T2
   auto X = [&](int ec) { return smem + (size_t)ec*PE; };
   auto T1 = [&](int ec) { return smem + (size_t)ec*PE + C::S_X; };
   auto T2 = [&](int ec) { return smem + (size_t)ec*PE + C::S_X + C::S_T1; };
   pdl_wait();
   // gather, lanes along the element-local dof index
   for (int it = tid; it < nel*NC*C::ND; it += NT)
   {
      const int e = it / (NC*C::ND), r = it - e*(NC*C::ND);
      const int c = r / C::ND, i = r - c*C::ND;
      X(e*NC + c)[i % D1D + DP*(i / D1D)] = x[(size_t)c*cstride + __ldg(map + (size_t)(eb + e)*C::ND + i)];
   }
   __syncthreads();
   for (int it = tid; it < nel*NC*DD; it += NT)              // x pencils (ec, dz, dy)
   {
      const int ec = it / DD, r = it - ec*DD;
      double in[D1D], out[Q1D];
#pragma unroll
      for (int d = 0; d < D1D; d++) { in[d] = X(ec)[d + DP*r]; }
      pencil_fwd<D1D,Q1D>(tab.B, in, out);
#pragma unroll
      for (int q = 0; q < Q1D; q++) { T1(ec)[q + QP*r] = out[q]; }
   }
   __syncthreads();
   for (int it = tid; it < nel*NC*D1D*Q1D; it += NT)         // y pencils (ec, dz, qx)
   {
      const int ec = it / (D1D*Q1D), r = it - ec*(D1D*Q1D);
      const int qx = r % Q1D, dz = r / Q1D;
      double in[D1D], out[Q1D];
#pragma unroll
      for (int d = 0; d < D1D; d++) { in[d] = T1(ec)[qx + QP*(d + D1D*dz)]; }
      pencil_fwd<D1D,Q1D>(tab.B, in, out);
#pragma unroll
      for (int q = 0; q < Q1D; q++) { T2(ec)[qx + Q1D*(q + Q1D*dz)] = out[q]; }
   }
   __syncthreads();
   double den[NC];
#pragma unroll
   for (int c = 0; c < NC; c++) { den[c] = 0.0; }
   for (int it = tid; it < nel*QQ; it += NT)                 // z pencils (e, column): forward, scale by D, back; all components
   {
      const int e = it / QQ, col = it - e*QQ;
      const double *d = Dq + (size_t)(eb + e)*C::NQ + col;
      double w[Q1D];
#pragma unroll
      for (int q = 0; q < Q1D; q++) { w[q] = __ldg(d + QQ*q); }
#pragma unroll
      for (int c = 0; c < NC; c++)
      {
         double *t2 = T2(e*NC + c) + col;
         double in[D1D], u[Q1D], o[D1D];
#pragma unroll
         for (int k = 0; k < D1D; k++) { in[k] = t2[QQ*k]; }
         pencil_fwd<D1D,Q1D>(tab.B, in, u);
#pragma unroll
         for (int q = 0; q < Q1D; q++)
         {
            const double dw = w[q]*u[q];
            if (WITH_DEN) { den[c] += dw*u[q]; }
            u[q] = dw;
         }
         pencil_bwd<D1D,Q1D>(tab.B, u, o);
#pragma unroll
         for (int k = 0; k < D1D; k++) { t2[QQ*k] = o[k]; }
      }
   }
   __syncthreads();
   for (int it = tid; it < nel*NC*D1D*Q1D; it += NT)         // y pencils back
   {
      const int ec = it / (D1D*Q1D), r = it - ec*(D1D*Q1D);
      const int qx = r % Q1D, dz = r / Q1D;
      double in[Q1D], o[D1D];
#pragma unroll
      for (int q = 0; q < Q1D; q++) { in[q] = T2(ec)[qx + Q1D*(q + Q1D*dz)]; }
      pencil_bwd<D1D,Q1D>(tab.B, in, o);
#pragma unroll
      for (int d = 0; d < D1D; d++) { T1(ec)[qx + QP*(d + D1D*dz)] = o[d]; }
   }
   __syncthreads();
   for (int it = tid; it < nel*NC*DD; it += NT)              // x pencils back
   {
      const int ec = it / DD, r = it - ec*DD;
      double in[Q1D], o[D1D];
#pragma unroll
      for (int q = 0; q < Q1D; q++) { in[q] = T1(ec)[q + QP*r]; }
      pencil_bwd<D1D,Q1D>(tab.B, in, o);
#pragma unroll
      for (int d = 0; d < D1D; d++) { X(ec)[d + DP*r] = o[d]; }
   }
   __syncthreads();
   for (int it = tid; it < nel*NC*C::ND; it += NT)           // scatter-add, lanes along the dof index
   {
      const int e = it / (NC*C::ND), r = it - e*(NC*C::ND);
      const int c = r / C::ND, i = r - c*C::ND;
      atomicAdd(y + (size_t)c*cstride + __ldg(map + (size_t)(eb + e)*C::ND + i), X(e*NC + c)[i % D1D + DP*(i / D1D)]);
   }
   if (WITH_DEN)
   {
      // deterministic block reduction (warp shuffle tree, warp partials in order); the PCG's finish kernel adds the CTAs
      __syncthreads();
      double *red = smem;
      constexpr int NW = NT/32;
#pragma unroll
      for (int c = 0; c < NC; c++)
      {
         double v = den[c];
         for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
         if ((tid & 31) == 0) { red[c*NW + (tid >> 5)] = v; }
      }
      __syncthreads();
      if (tid < NC)
      {
         double s = 0.0;
         for (int w = 0; w < NW; w++) { s += red[tid*NW + w]; }
         den_part[(size_t)blockIdx.x*NC + tid] = s;
      }
   }
}

} // namespace tuned
} // namespace lagb
