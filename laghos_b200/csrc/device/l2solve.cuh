// Direct solve with the L2 (thermodynamic) mass matrix.
//
// The reference solves M_e de = rhs with an unpreconditioned global CG (CG_EMass.Mult,
// laghos_solver.cpp:278-283 and :481).  M_e = sum_q rho0DetJ0w(q) phi_i(q) phi_j(q) is block
// diagonal (one NL x NL block per element, L2 space) and constant in time (the quadrature
// coefficient is the t = 0 mass, laghos_solver.cpp:232-249), so SURVEY.md 8f-1 replaces the ~12
// operator applications per solve by one product with the element inverses built once at setup:
//   l2inv_build  one CTA per element: assemble the block in shared memory from the 1D Bernstein
//                table and the quadrature coefficient, invert in place (Gauss-Jordan without
//                pivoting: the block is SPD), store [e][j][i];
//   l2inv_apply  y_e = M_e^-1 x_e, one thread per (element, row), lanes along the row index so
//                that every load of the inverse is contiguous; a pure HBM stream of NL^2 doubles
//                per element (Q3Q2: 5.8 KB against 12 x 1.7 KB of quadrature data for the CG).
// Runtime dimensions (all orders / both space dimensions share the two kernels).
#pragma once
#include "common.cuh"

namespace lagb {
namespace l2 {

// BL: [q + Q1D*l] 1D Bernstein table; D: [NE*NQ] quadrature coefficient; Minv: [NE][NL*NL]
__global__ void __launch_bounds__(256)
l2inv_build(const int dim, const int L1D, const int Q1D, const double *__restrict__ BL,
            const double *__restrict__ D, double *__restrict__ Minv)
{
   extern __shared__ double sm[];
   const int NL = (dim == 3) ? L1D*L1D*L1D : L1D*L1D;
   const int NQ = (dim == 3) ? Q1D*Q1D*Q1D : Q1D*Q1D;
   double *A = sm;                 // [NL][NL]
   double *colk = A + NL*NL;       // [NL]
   double *sBL = colk + NL;        // [Q1D*L1D]
   double *sD = sBL + Q1D*L1D;     // [NQ]
   const int e = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
   for (int i = tid; i < Q1D*L1D; i += nt) { sBL[i] = BL[i]; }
   for (int q = tid; q < NQ; q += nt) { sD[q] = D[(size_t)e*NQ + q]; }
   __syncthreads();
   // lower triangle (j <= i), mirrored
   for (int p = tid; p < NL*NL; p += nt)
   {
      const int i = p / NL, j = p - i*NL;
      if (j > i) { continue; }
      const int ix = i % L1D, iy = (i / L1D) % L1D, iz = i / (L1D*L1D);
      const int jx = j % L1D, jy = (j / L1D) % L1D, jz = j / (L1D*L1D);
      double acc = 0.0;
      const int QZ = (dim == 3) ? Q1D : 1;
      for (int qz = 0; qz < QZ; qz++)
      {
         const double bz = (dim == 3) ? sBL[qz + Q1D*iz]*sBL[qz + Q1D*jz] : 1.0;
         for (int qy = 0; qy < Q1D; qy++)
         {
            const double byz = bz*sBL[qy + Q1D*iy]*sBL[qy + Q1D*jy];
            const double *d = sD + Q1D*(qy + Q1D*qz);
            double row = 0.0;
            for (int qx = 0; qx < Q1D; qx++) { row += sBL[qx + Q1D*ix]*sBL[qx + Q1D*jx]*d[qx]; }
            acc += byz*row;
         }
      }
      A[i*NL + j] = acc; A[j*NL + i] = acc;
   }
   __syncthreads();
   // in-place Gauss-Jordan inversion, no pivoting
   for (int k = 0; k < NL; k++)
   {
      const double piv = 1.0/A[k*NL + k];
      __syncthreads();
      for (int i = tid; i < NL; i += nt) { colk[i] = A[i*NL + k]; }
      __syncthreads();
      for (int j = tid; j < NL; j += nt) { A[k*NL + j] = (j == k) ? piv : A[k*NL + j]*piv; }
      __syncthreads();
      for (int p = tid; p < NL*NL; p += nt)
      {
         const int i = p / NL, j = p - i*NL;
         if (i == k) { continue; }
         const double base = (j == k) ? 0.0 : A[p];
         A[p] = base - colk[i]*A[k*NL + j];
      }
      __syncthreads();
   }
   // symmetric up to round-off: store the average so that [j][i] == [i][j] exactly
   for (int p = tid; p < NL*NL; p += nt)
   {
      const int i = p / NL, j = p - i*NL;
      Minv[(size_t)e*NL*NL + p] = 0.5*(A[i*NL + j] + A[j*NL + i]);
   }
}

// Reference ComputeDensity (laghos_solver.cpp:542-563): per element rho_z = Mrho^-1 rhs with
// Mrho = sum_q wdet(q) phi_i phi_j (mass matrix on the current mesh) and rhs_i = sum_q rho0DetJ0w(q) phi_i
// (DensityIntegrator, laghos_assembly.cpp:26-41).  One CTA per element, same assembly and in-place
// Gauss-Jordan as l2inv_build; a diagnostics path (visualisation, -err).
__global__ void __launch_bounds__(256)
density_project(const int dim, const int L1D, const int Q1D, const double *__restrict__ BL,
                const double *__restrict__ wdet, const double *__restrict__ rho0DetJ0w, double *__restrict__ rho)
{
   extern __shared__ double sm[];
   const int NL = (dim == 3) ? L1D*L1D*L1D : L1D*L1D;
   const int NQ = (dim == 3) ? Q1D*Q1D*Q1D : Q1D*Q1D;
   double *A = sm, *colk = A + NL*NL, *sBL = colk + NL, *sD = sBL + Q1D*L1D, *sR = sD + NQ, *rhs = sR + NQ;
   const int e = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
   const int QZ = (dim == 3) ? Q1D : 1;
   for (int i = tid; i < Q1D*L1D; i += nt) { sBL[i] = BL[i]; }
   for (int q = tid; q < NQ; q += nt) { sD[q] = wdet[(size_t)e*NQ + q]; sR[q] = rho0DetJ0w[(size_t)e*NQ + q]; }
   __syncthreads();
   for (int p = tid; p < NL*NL + NL; p += nt)
   {
      const bool is_rhs = p >= NL*NL;
      const int i = is_rhs ? p - NL*NL : p / NL, j = is_rhs ? -1 : p - i*NL;
      if (!is_rhs && j > i) { continue; }
      const int ix = i % L1D, iy = (i / L1D) % L1D, iz = i / (L1D*L1D);
      const int jx = is_rhs ? 0 : j % L1D, jy = is_rhs ? 0 : (j / L1D) % L1D, jz = is_rhs ? 0 : j / (L1D*L1D);
      const double *coef = is_rhs ? sR : sD;
      double acc = 0.0;
      for (int qz = 0; qz < QZ; qz++)
      {
         const double bz = (dim == 3) ? sBL[qz + Q1D*iz]*(is_rhs ? 1.0 : sBL[qz + Q1D*jz]) : 1.0;
         for (int qy = 0; qy < Q1D; qy++)
         {
            const double byz = bz*sBL[qy + Q1D*iy]*(is_rhs ? 1.0 : sBL[qy + Q1D*jy]);
            const double *d = coef + Q1D*(qy + Q1D*qz);
            double row = 0.0;
            for (int qx = 0; qx < Q1D; qx++) { row += sBL[qx + Q1D*ix]*(is_rhs ? 1.0 : sBL[qx + Q1D*jx])*d[qx]; }
            acc += byz*row;
         }
      }
      if (is_rhs) { rhs[i] = acc; }
      else { A[i*NL + j] = acc; A[j*NL + i] = acc; }
   }
   __syncthreads();
   for (int k = 0; k < NL; k++)
   {
      const double piv = 1.0/A[k*NL + k];
      __syncthreads();
      for (int i = tid; i < NL; i += nt) { colk[i] = A[i*NL + k]; }
      __syncthreads();
      for (int j = tid; j < NL; j += nt) { A[k*NL + j] = (j == k) ? piv : A[k*NL + j]*piv; }
      __syncthreads();
      for (int p = tid; p < NL*NL; p += nt)
      {
         const int i = p / NL, j = p - i*NL;
         if (i == k) { continue; }
         const double base = (j == k) ? 0.0 : A[p];
         A[p] = base - colk[i]*A[k*NL + j];
      }
      __syncthreads();
   }
   for (int i = tid; i < NL; i += nt)
   {
      double r = 0.0;
      for (int j = 0; j < NL; j++) { r += A[i*NL + j]*rhs[j]; }
      rho[(size_t)e*NL + i] = r;
   }
}

// y[e][i] = sum_j Minv[e][j][i] x[e][j]; EPB elements per CTA, NLC = compile-time NL (0: runtime)
template<int NLC, int EPB>
__global__ void __launch_bounds__(NLC > 0 ? ((NLC*EPB + 31)/32)*32 : 256)
l2inv_apply(const int NE, const int NLr, const double *__restrict__ Minv, const double *__restrict__ x,
            double *__restrict__ y)
{
   const int NL = (NLC > 0) ? NLC : NLr;
   extern __shared__ double sx[];   // [EPB][NL]
   const int tid = threadIdx.x;
   const int e0 = blockIdx.x*EPB;
   const int nel = min(EPB, NE - e0);
   for (int i = tid; i < nel*NL; i += blockDim.x) { sx[i] = x[(size_t)e0*NL + i]; }
   __syncthreads();
   for (int t = tid; t < nel*NL; t += blockDim.x)
   {
      const int el = t / NL, i = t - el*NL;
      const double *m = Minv + (size_t)(e0 + el)*NL*NL + i;
      const double *xe = sx + el*NL;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int j = 0;
      if (NLC > 0)
      {
         // all loads of the row sweep are independent: issue them before the first use
         double mv[NLC > 0 ? NLC : 1];
#pragma unroll
         for (int jj = 0; jj < NLC; jj++) { mv[jj] = __ldcs(m + (size_t)jj*NLC); }
#pragma unroll
         for (int jj = 0; jj < NLC; jj++)
         {
            if ((jj & 3) == 0) { a0 += mv[jj]*xe[jj]; }
            else if ((jj & 3) == 1) { a1 += mv[jj]*xe[jj]; }
            else if ((jj & 3) == 2) { a2 += mv[jj]*xe[jj]; }
            else { a3 += mv[jj]*xe[jj]; }
         }
      }
      else
      {
         for (; j + 3 < NL; j += 4)
         {
            const double m0 = __ldcs(m + (size_t)j*NL), m1 = __ldcs(m + (size_t)(j + 1)*NL);
            const double m2 = __ldcs(m + (size_t)(j + 2)*NL), m3 = __ldcs(m + (size_t)(j + 3)*NL);
            a0 += m0*xe[j]; a1 += m1*xe[j + 1]; a2 += m2*xe[j + 2]; a3 += m3*xe[j + 3];
         }
         for (; j < NL; j++) { a0 += __ldcs(m + (size_t)j*NL)*xe[j]; }
      }
      y[(size_t)(e0 + el)*NL + i] = (a0 + a1) + (a2 + a3);
   }
}

} // namespace l2
} // namespace lagb
