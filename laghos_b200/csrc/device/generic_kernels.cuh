// Generic kernels: one thread per element, every (DIM, D1D, Q1D) the reference
// instantiates (laghos_assembly.cpp:536-548, laghos_solver.cpp:1387-1396) plus 3D
// (6,10).  These are the path for 2D problems (tiny meshes, gate config 1) and the
// in-library cross-check variant (kernel_variant = 1) for the tuned 3D kernels.
// Element-local arrays live in local memory; performance is not the point here.
#pragma once
#include "common.cuh"

namespace lagb {
namespace generic {

template<int DIM, int D1D, int Q1D>
struct Dims
{
   static constexpr int L1D = D1D - 1;
   static constexpr int DZ = (DIM == 3) ? D1D : 1;
   static constexpr int QZ = (DIM == 3) ? Q1D : 1;
   static constexpr int LZ = (DIM == 3) ? L1D : 1;
   static constexpr int ND = D1D*D1D*DZ, NQ = Q1D*Q1D*QZ, NL = L1D*L1D*LZ;
};

// dofs -> quadrature values with table T (Q1D x N1D)
template<int DIM, int N1D, int Q1D>
__device__ inline void interp(const double *T, const double *in, double *out, double *t0, double *t1)
{
   constexpr int NZ = (DIM == 3) ? N1D : 1;
   contract<N1D, Q1D, N1D, N1D, NZ, 0, false>(T, in, t0);
   if (DIM == 3)
   {
      contract<N1D, Q1D, Q1D, N1D, NZ, 1, false>(T, t0, t1);
      contract<N1D, Q1D, Q1D, Q1D, NZ, 2, false>(T, t1, out);
   }
   else { contract<N1D, Q1D, Q1D, N1D, NZ, 1, false>(T, t0, out); }
}
// quadrature -> dofs (transpose), tables Tx,Ty,Tz each (Q1D x N1D)
template<int DIM, int N1D, int Q1D>
__device__ inline void interp_t(const double *Tx, const double *Ty, const double *Tz,
                                const double *in, double *out, double *t0, double *t1)
{
   constexpr int QZ = (DIM == 3) ? Q1D : 1;
   contract<Q1D, N1D, Q1D, Q1D, QZ, 0, true>(Tx, in, t0);
   if (DIM == 3)
   {
      contract<Q1D, N1D, N1D, Q1D, QZ, 1, true>(Ty, t0, t1);
      contract<Q1D, N1D, N1D, N1D, QZ, 2, true>(Tz, t1, out);
   }
   else { contract<Q1D, N1D, N1D, Q1D, QZ, 1, true>(Ty, t0, out); }
}

template<int DIM, int D1D, int Q1D>
__global__ void mass_h1(const __grid_constant__ DevTables<D1D,Q1D> tab, int NE,
                        const int *__restrict__ map, const double *__restrict__ D,
                        const double *__restrict__ x, double *__restrict__ y)
{
   using Dm = Dims<DIM,D1D,Q1D>;
   for (int e = blockIdx.x*blockDim.x + threadIdx.x; e < NE; e += gridDim.x*blockDim.x)
   {
      double X[Dm::ND], QQ[Dm::NQ], t0[Q1D*D1D*Dm::QZ], t1[Q1D*Q1D*Dm::DZ];
      const int *m = map + (size_t)e*Dm::ND;
      for (int i = 0; i < Dm::ND; i++) { X[i] = x[m[i]]; }
      interp<DIM,D1D,Q1D>(tab.B, X, QQ, t0, t1);
      const double *d = D + (size_t)e*Dm::NQ;
      for (int q = 0; q < Dm::NQ; q++) { QQ[q] *= d[q]; }
      interp_t<DIM,D1D,Q1D>(tab.B, tab.B, tab.B, QQ, X, t0, t1);
      for (int i = 0; i < Dm::ND; i++) { atomicAdd(&y[m[i]], X[i]); }
   }
}

template<int DIM, int D1D, int Q1D>
__global__ void mass_h1_diag(const __grid_constant__ DevTables<D1D,Q1D> tab, int NE,
                             const int *__restrict__ map, const double *__restrict__ D,
                             double *__restrict__ diag)
{
   using Dm = Dims<DIM,D1D,Q1D>;
   double B2[Q1D*D1D];
   for (int i = 0; i < Q1D*D1D; i++) { B2[i] = tab.B[i]*tab.B[i]; }
   for (int e = blockIdx.x*blockDim.x + threadIdx.x; e < NE; e += gridDim.x*blockDim.x)
   {
      double Y[Dm::ND], t0[Q1D*D1D*Dm::QZ], t1[Q1D*Q1D*Dm::DZ];
      const int *m = map + (size_t)e*Dm::ND;
      interp_t<DIM,D1D,Q1D>(B2, B2, B2, D + (size_t)e*Dm::NQ, Y, t0, t1);
      for (int i = 0; i < Dm::ND; i++) { atomicAdd(&diag[m[i]], Y[i]); }
   }
}

template<int DIM, int D1D, int Q1D>
__global__ void mass_l2(const __grid_constant__ DevTables<D1D,Q1D> tab, int NE,
                        const double *__restrict__ D, const double *__restrict__ x, double *__restrict__ y)
{
   using Dm = Dims<DIM,D1D,Q1D>;
   constexpr int L1D = Dm::L1D;
   for (int e = blockIdx.x*blockDim.x + threadIdx.x; e < NE; e += gridDim.x*blockDim.x)
   {
      double X[Dm::NL], QQ[Dm::NQ], t0[Q1D*D1D*Dm::QZ], t1[Q1D*Q1D*Dm::DZ];
      for (int i = 0; i < Dm::NL; i++) { X[i] = x[(size_t)e*Dm::NL + i]; }
      interp<DIM,L1D,Q1D>(tab.BL, X, QQ, t0, t1);
      const double *d = D + (size_t)e*Dm::NQ;
      for (int q = 0; q < Dm::NQ; q++) { QQ[q] *= d[q]; }
      interp_t<DIM,L1D,Q1D>(tab.BL, tab.BL, tab.BL, QQ, X, t0, t1);
      for (int i = 0; i < Dm::NL; i++) { y[(size_t)e*Dm::NL + i] = X[i]; }
   }
}

// reference ForceMult2D/3D + H1R->MultTranspose
template<int DIM, int D1D, int Q1D>
__global__ void force_mult(const __grid_constant__ DevTables<D1D,Q1D> tab, int NE, int64_t ndofs,
                           const int *__restrict__ map, const double *__restrict__ sJit,
                           const double *__restrict__ x, double *__restrict__ y)
{
   using Dm = Dims<DIM,D1D,Q1D>;
   constexpr int L1D = Dm::L1D;
   const double eps2 = DBL_EPSILON*DBL_EPSILON;
   const size_t NEQ = (size_t)NE*Dm::NQ;
   for (int e = blockIdx.x*blockDim.x + threadIdx.x; e < NE; e += gridDim.x*blockDim.x)
   {
      double E[Dm::NL], QQQ[Dm::NQ], QQg[Dm::NQ], acc[Dm::ND], part[Dm::ND];
      double t0[Q1D*D1D*Dm::QZ], t1[Q1D*Q1D*Dm::DZ];
      const int *m = map + (size_t)e*Dm::ND;
      for (int i = 0; i < Dm::NL; i++) { E[i] = x[(size_t)e*Dm::NL + i]; }
      interp<DIM,L1D,Q1D>(tab.BL, E, QQQ, t0, t1);
      for (int c = 0; c < DIM; c++)
      {
         for (int g = 0; g < DIM; g++)
         {
            const double *s = sJit + (size_t)e*Dm::NQ + NEQ*(g + DIM*c);
            for (int q = 0; q < Dm::NQ; q++) { QQg[q] = QQQ[q]*s[q]; }
            interp_t<DIM,D1D,Q1D>(g == 0 ? tab.G : tab.B, g == 1 ? tab.G : tab.B,
                                  g == 2 ? tab.G : tab.B, QQg, part, t0, t1);
            if (g == 0) { for (int i = 0; i < Dm::ND; i++) { acc[i] = part[i]; } }
            else { for (int i = 0; i < Dm::ND; i++) { acc[i] += part[i]; } }
         }
         for (int i = 0; i < Dm::ND; i++)
         {
            double v = acc[i];
            if (fabs(v) < eps2) { v = 0.0; }
            atomicAdd(&y[(size_t)c*ndofs + m[i]], v);
         }
      }
   }
}

// reference H1R->Mult + ForceMultTranspose2D/3D
template<int DIM, int D1D, int Q1D>
__global__ void force_mult_t(const __grid_constant__ DevTables<D1D,Q1D> tab, int NE, int64_t ndofs,
                             const int *__restrict__ map, const double *__restrict__ sJit,
                             const double *__restrict__ v, double *__restrict__ eout)
{
   using Dm = Dims<DIM,D1D,Q1D>;
   constexpr int L1D = Dm::L1D;
   const size_t NEQ = (size_t)NE*Dm::NQ;
   for (int e = blockIdx.x*blockDim.x + threadIdx.x; e < NE; e += gridDim.x*blockDim.x)
   {
      double V[Dm::ND], QQQ[Dm::NQ], dq[Dm::NQ], sum[Dm::NQ];
      double t0[Q1D*D1D*Dm::QZ], t1[Q1D*Q1D*Dm::DZ];
      const int *m = map + (size_t)e*Dm::ND;
      for (int q = 0; q < Dm::NQ; q++) { QQQ[q] = 0.0; }
      for (int c = 0; c < DIM; c++)
      {
         for (int i = 0; i < Dm::ND; i++) { V[i] = v[(size_t)c*ndofs + m[i]]; }
         for (int g = 0; g < DIM; g++)
         {
            // gradient component g: G along axis g, B along the others
            contract<D1D, Q1D, D1D, D1D, Dm::DZ, 0, false>(g == 0 ? tab.G : tab.B, V, t0);
            if (DIM == 3)
            {
               contract<D1D, Q1D, Q1D, D1D, Dm::DZ, 1, false>(g == 1 ? tab.G : tab.B, t0, t1);
               contract<D1D, Q1D, Q1D, Q1D, Dm::DZ, 2, false>(g == 2 ? tab.G : tab.B, t1, dq);
            }
            else { contract<D1D, Q1D, Q1D, D1D, Dm::DZ, 1, false>(g == 1 ? tab.G : tab.B, t0, dq); }
            const double *s = sJit + (size_t)e*Dm::NQ + NEQ*(g + DIM*c);
            if (g == 0) { for (int q = 0; q < Dm::NQ; q++) { sum[q] = dq[q]*s[q]; } }
            else { for (int q = 0; q < Dm::NQ; q++) { sum[q] += dq[q]*s[q]; } }
         }
         for (int q = 0; q < Dm::NQ; q++) { QQQ[q] += sum[q]; }
      }
      double out[Dm::NL];
      interp_t<DIM,L1D,Q1D>(tab.BL, tab.BL, tab.BL, QQQ, out, t0, t1);
      for (int i = 0; i < Dm::NL; i++) { eout[(size_t)e*Dm::NL + i] = out[i]; }
   }
}

// Gradients of NF vector fields (each DIM components) and the value of one L2 field
// at the points of one z-plane, evaluated point by point from z-contracted planes.
// Used by qupdate, rho0detj0 and taylor_source below.
template<int DIM, int D1D, int Q1D>
struct PlaneEval
{
   using Dm = Dims<DIM,D1D,Q1D>;
   static constexpr int L1D = Dm::L1D;
   static constexpr int DD = D1D*D1D, LL = L1D*L1D;

   // z-contract field X[ND] at plane qz: WB = sum_dz B(qz,dz) X, WG = sum_dz G(qz,dz) X
   __device__ static inline void zplane(const DevTables<D1D,Q1D> &tab, int qz, const double *X, double *WB, double *WG)
   {
      if (DIM == 2)
      {
         for (int i = 0; i < DD; i++) { WB[i] = X[i]; }
         return;
      }
      for (int i = 0; i < DD; i++)
      {
         double b = 0.0, g = 0.0;
#pragma unroll
         for (int dz = 0; dz < D1D; dz++)
         {
            b += tab.B[qz + Q1D*dz]*X[i + DD*dz];
            g += tab.G[qz + Q1D*dz]*X[i + DD*dz];
         }
         WB[i] = b; WG[i] = g;
      }
   }
   __device__ static inline void zplane_l2(const DevTables<D1D,Q1D> &tab, int qz, const double *E, double *WE)
   {
      if (DIM == 2)
      {
         for (int i = 0; i < LL; i++) { WE[i] = E[i]; }
         return;
      }
      for (int i = 0; i < LL; i++)
      {
         double b = 0.0;
#pragma unroll
         for (int lz = 0; lz < L1D; lz++) { b += tab.BL[qz + Q1D*lz]*E[i + LL*lz]; }
         WE[i] = b;
      }
   }
   // gradient (d/dxi_0, d/dxi_1[, d/dxi_2]) and value at point (qx,qy) of the plane
   __device__ static inline void point(const DevTables<D1D,Q1D> &tab, int qx, int qy,
                                       const double *WB, const double *WG, double *grad, double &val)
   {
      double gx = 0.0, gy = 0.0, gz = 0.0, vv = 0.0;
      for (int dy = 0; dy < D1D; dy++)
      {
         double bx = 0.0, dx_ = 0.0, bz = 0.0;
#pragma unroll
         for (int dx = 0; dx < D1D; dx++)
         {
            const double w = WB[dx + D1D*dy];
            bx += tab.B[qx + Q1D*dx]*w;
            dx_ += tab.G[qx + Q1D*dx]*w;
            if (DIM == 3) { bz += tab.B[qx + Q1D*dx]*WG[dx + D1D*dy]; }
         }
         gx += tab.B[qy + Q1D*dy]*dx_;
         gy += tab.G[qy + Q1D*dy]*bx;
         vv += tab.B[qy + Q1D*dy]*bx;
         if (DIM == 3) { gz += tab.B[qy + Q1D*dy]*bz; }
      }
      grad[0] = gx; grad[1] = gy; if (DIM == 3) { grad[2] = gz; }
      val = vv;
   }
   __device__ static inline double point_l2(const DevTables<D1D,Q1D> &tab, int qx, int qy, const double *WE)
   {
      double vv = 0.0;
      for (int ly = 0; ly < L1D; ly++)
      {
         double bx = 0.0;
#pragma unroll
         for (int lx = 0; lx < L1D; lx++) { bx += tab.BL[qx + Q1D*lx]*WE[lx + L1D*ly]; }
         vv += tab.BL[qy + Q1D*ly]*bx;
      }
      return vv;
   }
};

// reference QUpdate::UpdateQuadratureData (laghos_solver.cpp:1354-1411) fused:
// restriction + interpolation + point physics + per-thread dt minimum.
template<int DIM, int D1D, int Q1D>
__global__ void qupdate(const __grid_constant__ DevTables<D1D,Q1D> tab, int NE, int64_t ndofs,
                        const int *__restrict__ map, const double *__restrict__ S,
                        const double *__restrict__ rho0DetJ0w, const double *__restrict__ Jac0inv,
                        const double *__restrict__ gamma, const double *__restrict__ qweights,
                        QPointParams prm, double *__restrict__ sJit, double *__restrict__ dt_block_min)
{
   using Dm = Dims<DIM,D1D,Q1D>;
   using PE = PlaneEval<DIM,D1D,Q1D>;
   constexpr int DIM2 = DIM*DIM;
   const double *x = S, *v = S + DIM*ndofs, *en = S + 2*DIM*ndofs;
   const size_t NEQ = (size_t)NE*Dm::NQ;
   double dt_min = prm.dt_in;
   for (int e = blockIdx.x*blockDim.x + threadIdx.x; e < NE; e += gridDim.x*blockDim.x)
   {
      double X[DIM][Dm::ND], V[DIM][Dm::ND], E[Dm::NL];
      double WB[2*DIM][PE::DD], WG[2*DIM][PE::DD], WE[PE::LL > 0 ? PE::LL : 1];
      const int *m = map + (size_t)e*Dm::ND;
      for (int c = 0; c < DIM; c++)
         for (int i = 0; i < Dm::ND; i++)
         {
            X[c][i] = x[(size_t)c*ndofs + m[i]];
            V[c][i] = v[(size_t)c*ndofs + m[i]];
         }
      for (int i = 0; i < Dm::NL; i++) { E[i] = en[(size_t)e*Dm::NL + i]; }
      const double gam = gamma[e];
      for (int qz = 0; qz < Dm::QZ; qz++)
      {
         for (int c = 0; c < DIM; c++)
         {
            PE::zplane(tab, qz, X[c], WB[c], WG[c]);
            PE::zplane(tab, qz, V[c], WB[DIM + c], WG[DIM + c]);
         }
         PE::zplane_l2(tab, qz, E, WE);
         for (int qy = 0; qy < Q1D; qy++)
            for (int qx = 0; qx < Q1D; qx++)
            {
               const int q = qx + Q1D*(qy + Q1D*qz);
               double J[DIM2], dV[DIM2], g[3], val;
               for (int c = 0; c < DIM; c++)
               {
                  PE::point(tab, qx, qy, WB[c], WG[c], g, val);
                  for (int d = 0; d < DIM; d++) { J[c + DIM*d] = g[d]; }
                  PE::point(tab, qx, qy, WB[DIM + c], WG[DIM + c], g, val);
                  for (int d = 0; d < DIM; d++) { dV[c + DIM*d] = g[d]; }
               }
               const double e_q = PE::point_l2(tab, qx, qy, WE);
               const size_t eq = (size_t)e*Dm::NQ + q;
               double J0[DIM2], sJ[DIM2];
               for (int k = 0; k < DIM2; k++) { J0[k] = Jac0inv[eq*DIM2 + k]; }
               const double dtq = qpoint<DIM>(J, dV, e_q, rho0DetJ0w[eq], J0, gam, qweights[q], 1.0/qweights[q], prm, sJ);
               dt_min = fmin(dt_min, dtq);
               for (int vd = 0; vd < DIM; vd++)
                  for (int gd = 0; gd < DIM; gd++) { sJit[eq + NEQ*(gd + vd*DIM)] = sJ[vd + gd*DIM]; }
            }
      }
   }
   // block minimum (fmin is exact and order-independent: deterministic)
   __shared__ double smin[32];
   for (int o = 16; o > 0; o >>= 1) { dt_min = fmin(dt_min, __shfl_xor_sync(0xffffffffu, dt_min, o)); }
   if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = dt_min; }
   __syncthreads();
   if (threadIdx.x == 0)
   {
      double m = smin[0];
      for (int w = 1; w < (blockDim.x + 31)/32; w++) { m = fmin(m, smin[w]); }
      atomic_min_nonneg(dt_block_min, m);   // dt_block_min: the context's running dt estimate
   }
}

// reference Rho0DetJ0Vol (laghos_solver.cpp:1170-1261): rho0DetJ0w, Jac0inv,
// mass coefficient D = w*rho0_q*detJ0, and per-element volume (summed on the host side
// of the C ABI in a fixed order).
template<int DIM, int D1D, int Q1D>
__global__ void rho0detj0(const __grid_constant__ DevTables<D1D,Q1D> tab, int NE, int64_t ndofs,
                          const int *__restrict__ map, const double *__restrict__ x0,
                          const double *__restrict__ rho0_gf, const double *__restrict__ rho0_q,
                          const double *__restrict__ qweights,
                          double *__restrict__ rho0DetJ0w, double *__restrict__ Jac0inv,
                          double *__restrict__ massD, double *__restrict__ elem_vol)
{
   using Dm = Dims<DIM,D1D,Q1D>;
   using PE = PlaneEval<DIM,D1D,Q1D>;
   constexpr int DIM2 = DIM*DIM;
   for (int e = blockIdx.x*blockDim.x + threadIdx.x; e < NE; e += gridDim.x*blockDim.x)
   {
      double X[DIM][Dm::ND], E[Dm::NL];
      double WB[DIM][PE::DD], WG[DIM][PE::DD], WE[PE::LL > 0 ? PE::LL : 1];
      const int *m = map + (size_t)e*Dm::ND;
      for (int c = 0; c < DIM; c++)
         for (int i = 0; i < Dm::ND; i++) { X[c][i] = x0[(size_t)c*ndofs + m[i]]; }
      for (int i = 0; i < Dm::NL; i++) { E[i] = rho0_gf[(size_t)e*Dm::NL + i]; }
      double vol = 0.0;
      for (int qz = 0; qz < Dm::QZ; qz++)
      {
         for (int c = 0; c < DIM; c++) { PE::zplane(tab, qz, X[c], WB[c], WG[c]); }
         PE::zplane_l2(tab, qz, E, WE);
         for (int qy = 0; qy < Q1D; qy++)
            for (int qx = 0; qx < Q1D; qx++)
            {
               const int q = qx + Q1D*(qy + Q1D*qz);
               double J[DIM2], g[3], val;
               for (int c = 0; c < DIM; c++)
               {
                  PE::point(tab, qx, qy, WB[c], WG[c], g, val);
                  for (int d = 0; d < DIM; d++) { J[c + DIM*d] = g[d]; }
               }
               const double R = PE::point_l2(tab, qx, qy, WE);
               const size_t eq = (size_t)e*Dm::NQ + q;
               const double W = qweights[q];
               double *inv = Jac0inv + eq*DIM2;
               double det;
               if (DIM == 2)
               {
                  det = J[0]*J[3] - J[1]*J[2];
                  const double r = 1.0/det;
                  inv[0] = J[3]*r; inv[1] = -J[1]*r; inv[2] = -J[2]*r; inv[3] = J[0]*r;
               }
               else
               {
                  det = J[0]*(J[4]*J[8] - J[5]*J[7]) + J[3]*(J[2]*J[7] - J[1]*J[8]) + J[6]*(J[1]*J[5] - J[2]*J[4]);
                  const double r = 1.0/det;
                  const double J11 = J[0], J21 = J[1], J31 = J[2], J12 = J[3], J22 = J[4], J32 = J[5];
                  const double J13 = J[6], J23 = J[7], J33 = J[8];
                  // placement as the reference writes it (laghos_solver.cpp:1243-1251):
                  // the transpose of the textbook inverse, kept for parity.
                  inv[0] = r*((J22*J33) - (J23*J32));
                  inv[1] = r*((J32*J13) - (J33*J12));
                  inv[2] = r*((J12*J23) - (J13*J22));
                  inv[3] = r*((J23*J31) - (J21*J33));
                  inv[4] = r*((J33*J11) - (J31*J13));
                  inv[5] = r*((J13*J21) - (J11*J23));
                  inv[6] = r*((J21*J32) - (J22*J31));
                  inv[7] = r*((J31*J12) - (J32*J11));
                  inv[8] = r*((J11*J22) - (J12*J21));
               }
               rho0DetJ0w[eq] = W*R*det;
               massD[eq] = W*(rho0_q ? rho0_q[eq] : R)*det;
               vol += W*det;
            }
      }
      elem_vol[e] = vol;
   }
}

// w(q) detJ(q) on the CURRENT mesh: the coefficient of MassIntegrator in ComputeDensity
// (reference laghos_solver.cpp:542-563); a diagnostics path (visualisation, -err), not timed.
template<int DIM, int D1D, int Q1D>
__global__ void detj_w(const __grid_constant__ DevTables<D1D,Q1D> tab, int NE, int64_t ndofs,
                       const int *__restrict__ map, const double *__restrict__ x,
                       const double *__restrict__ qweights, double *__restrict__ out)
{
   using Dm = Dims<DIM,D1D,Q1D>;
   using PE = PlaneEval<DIM,D1D,Q1D>;
   for (int e = blockIdx.x*blockDim.x + threadIdx.x; e < NE; e += gridDim.x*blockDim.x)
   {
      double X[DIM][Dm::ND];
      double WB[DIM][PE::DD], WG[DIM][PE::DD];
      const int *m = map + (size_t)e*Dm::ND;
      for (int c = 0; c < DIM; c++)
         for (int i = 0; i < Dm::ND; i++) { X[c][i] = x[(size_t)c*ndofs + m[i]]; }
      for (int qz = 0; qz < Dm::QZ; qz++)
      {
         for (int c = 0; c < DIM; c++) { PE::zplane(tab, qz, X[c], WB[c], WG[c]); }
         for (int qy = 0; qy < Q1D; qy++)
            for (int qx = 0; qx < Q1D; qx++)
            {
               const int q = qx + Q1D*(qy + Q1D*qz);
               double J[DIM*DIM], g[3], val;
               for (int c = 0; c < DIM; c++)
               {
                  PE::point(tab, qx, qy, WB[c], WG[c], g, val);
                  for (int d = 0; d < DIM; d++) { J[c + DIM*d] = g[d]; }
               }
               const double det = (DIM == 2) ? J[0]*J[3] - J[1]*J[2]
                                  : J[0]*(J[4]*J[DIM*DIM-1] - J[5]*J[7]) + J[3]*(J[2]*J[7] - J[1]*J[DIM*DIM-1]) + J[6]*(J[1]*J[5] - J[2]*J[4]);
               out[(size_t)e*Dm::NQ + q] = qweights[q]*det;
            }
      }
   }
}

// 2D Taylor-Green source (reference laghos_solver.cpp:455-465, laghos_solver.hpp:208-218)
template<int DIM, int D1D, int Q1D>
__global__ void taylor_source(const __grid_constant__ DevTables<D1D,Q1D> tab, int NE, int64_t ndofs,
                              const int *__restrict__ map, const double *__restrict__ x,
                              const double *__restrict__ qweights, double *__restrict__ esrc)
{
   using Dm = Dims<DIM,D1D,Q1D>;
   using PE = PlaneEval<DIM,D1D,Q1D>;
   constexpr int L1D = Dm::L1D;
   for (int e = blockIdx.x*blockDim.x + threadIdx.x; e < NE; e += gridDim.x*blockDim.x)
   {
      double X[DIM][Dm::ND], f[Dm::NQ], out[Dm::NL];
      double WB[DIM][PE::DD], WG[DIM][PE::DD];
      double t0[Q1D*D1D*Dm::QZ], t1[Q1D*Q1D*Dm::DZ];
      const int *m = map + (size_t)e*Dm::ND;
      for (int c = 0; c < DIM; c++)
         for (int i = 0; i < Dm::ND; i++) { X[c][i] = x[(size_t)c*ndofs + m[i]]; }
      for (int qz = 0; qz < Dm::QZ; qz++)
      {
         for (int c = 0; c < DIM; c++) { PE::zplane(tab, qz, X[c], WB[c], WG[c]); }
         for (int qy = 0; qy < Q1D; qy++)
            for (int qx = 0; qx < Q1D; qx++)
            {
               const int q = qx + Q1D*(qy + Q1D*qz);
               double J[DIM*DIM], g[3], xq[3] = {0, 0, 0};
               for (int c = 0; c < DIM; c++)
               {
                  PE::point(tab, qx, qy, WB[c], WG[c], g, xq[c]);
                  for (int d = 0; d < DIM; d++) { J[c + DIM*d] = g[d]; }
               }
               const double det = (DIM == 2) ? J[0]*J[3] - J[1]*J[2]
                                  : J[0]*(J[4]*J[DIM*DIM-1] - J[5]*J[7]) + J[3]*(J[2]*J[7] - J[1]*J[DIM*DIM-1]) + J[6]*(J[1]*J[5] - J[2]*J[4]);
               const double fx = 3.0/8.0*M_PI*(cos(3.0*M_PI*xq[0])*cos(M_PI*xq[1]) -
                                               cos(M_PI*xq[0])*cos(3.0*M_PI*xq[1]));
               f[q] = qweights[q]*det*fx;
            }
      }
      interp_t<DIM,L1D,Q1D>(tab.BL, tab.BL, tab.BL, f, out, t0, t1);
      for (int i = 0; i < Dm::NL; i++) { esrc[(size_t)e*Dm::NL + i] = out[i]; }
   }
}

} // namespace generic
} // namespace lagb
