// Tuned 3D kernels for QUpdate, Force, Force^T and the L2 mass apply.
//
//   qupdate3d   reference QUpdate::UpdateQuadratureData (laghos_solver.cpp:1354-1411):
//               H1R->Mult + q1->Derivatives (x and v) + q2->Values (e) + QKernel fused;
//               the q_dx / q_dv / q_e / q_dt_est temporaries of the reference never exist.
//   force3d     reference ForcePAOperator::Mult (laghos_assembly.cpp:557-565):
//               L2R->Mult + ForceMult3D (:296-514) + H1R->MultTranspose fused.
//   forcet3d    reference ForcePAOperator::MultTranspose (:965-973): H1R->Mult +
//               ForceMultTranspose3D (:715-924) + L2R->MultTranspose fused.
//   massl2_3d   reference MassPAOperator(L2)::Mult (laghos_solver.cpp:179; MFEM PA mass).
//
// Structure ("pencil" sum factorisation): a CTA owns NB elements; every 1D contraction
// stage is a flat list of pencils (one per line of the tensor along the contracted axis)
// distributed over all threads.  A pencil loads its N inputs from shared memory once and
// produces all M outputs in registers, so the 1D tables are compile-time indexed operands
// from the kernel-parameter constant bank (no table traffic) and shared memory sees
// (N + M) accesses per N*M FMAs.  Quadrature data (stressJinvT, Jac0inv, rho0DetJ0w, D) is
// read or written exactly once, q-contiguous (coalesced); L-vector gathers / scatter-adds
// run with lanes along the element-local dof index.
#pragma once
#include "common.cuh"
#include <cfloat>

namespace lagb {
namespace tuned {

// out[q] = sum_d T[q + Q*d]*in[d]   (dofs -> quadrature), compile-time table indices
template<int N1D, int Q1D>
__device__ __forceinline__ void pencil_fwd(const double *T, const double (&in)[N1D], double (&out)[Q1D])
{
#pragma unroll
   for (int q = 0; q < Q1D; q++)
   {
      double u = 0.0;
#pragma unroll
      for (int d = 0; d < N1D; d++) { u += T[q + Q1D*d]*in[d]; }
      out[q] = u;
   }
}
// out[d] = sum_q T[q + Q*d]*in[q]   (quadrature -> dofs)
template<int N1D, int Q1D>
__device__ __forceinline__ void pencil_bwd(const double *T, const double (&in)[Q1D], double (&out)[N1D])
{
#pragma unroll
   for (int d = 0; d < N1D; d++)
   {
      double u = 0.0;
#pragma unroll
      for (int q = 0; q < Q1D; q++) { u += T[q + Q1D*d]*in[q]; }
      out[d] = u;
   }
}

// 16-byte asynchronous global->shared copy (LDGSTS): no register staging, completes in the
// background while the gather and the x/y pencils run
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
   const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all()
{
   asm volatile("cp.async.commit_group;" ::: "memory");
   asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Bulk asynchronous copy (TMA, 1D): one thread hands a contiguous, 16-byte aligned slab to the copy engine,
// completion is counted in bytes on a shared-memory mbarrier (SASS: UBLKCP + SYNCS).
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
   const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count) : "memory");
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned int bytes)
{
   const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned int bytes, uint64_t *bar)
{
   const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
   const unsigned int b = (unsigned int)__cvta_generic_to_shared(bar);
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                :: "r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned int parity)
{
   const unsigned int a = (unsigned int)__cvta_generic_to_shared(bar);
   asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@!p bra WAIT_%=;\n\t}" :: "r"(a), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------
// L2 (Bernstein) dofs -> values at quadrature points for NB elements.
// Es[e][NL] -> Eq[e][NQ]; scratch t1[e][L*L*Q], t2[e][L*Q*Q].  Ends with a barrier.
// ---------------------------------------------------------------------------
template<int L1D, int Q1D>
__device__ __forceinline__ void l2_values(const double *BL, int nel, const double *Es, int sE,
                                          double *t1, int s1, double *t2, int s2, double *Eq, int sQ,
                                          int tid, int nthr)
{
   // rows of Es and t1 are padded to odd strides (LP, QP): lanes along the row index hit distinct banks
   constexpr int QQ = Q1D*Q1D, LL = L1D*L1D, QP = Q1D | 1, LP = L1D | 1;
   for (int it = tid; it < nel*LL; it += nthr)            // x pencils (lz, ly)
   {
      const int e = it / LL, r = it - e*LL;
      double in[L1D], out[Q1D];
#pragma unroll
      for (int l = 0; l < L1D; l++) { in[l] = Es[e*sE + l + LP*r]; }
      pencil_fwd<L1D,Q1D>(BL, in, out);
#pragma unroll
      for (int q = 0; q < Q1D; q++) { t1[e*s1 + q + QP*r] = out[q]; }     // [lz][ly][qx]
   }
   __syncthreads();
   for (int it = tid; it < nel*L1D*Q1D; it += nthr)       // y pencils (lz, qx)
   {
      const int e = it / (L1D*Q1D), r = it - e*(L1D*Q1D);
      const int qx = r % Q1D, lz = r / Q1D;
      double in[L1D], out[Q1D];
#pragma unroll
      for (int l = 0; l < L1D; l++) { in[l] = t1[e*s1 + qx + QP*(l + L1D*lz)]; }
      pencil_fwd<L1D,Q1D>(BL, in, out);
#pragma unroll
      for (int q = 0; q < Q1D; q++) { t2[e*s2 + qx + Q1D*(q + Q1D*lz)] = out[q]; }   // [lz][qy][qx]
   }
   __syncthreads();
   for (int it = tid; it < nel*QQ; it += nthr)            // z pencils (column)
   {
      const int e = it / QQ, col = it - e*QQ;
      double in[L1D], out[Q1D];
#pragma unroll
      for (int l = 0; l < L1D; l++) { in[l] = t2[e*s2 + col + QQ*l]; }
      pencil_fwd<L1D,Q1D>(BL, in, out);
#pragma unroll
      for (int q = 0; q < Q1D; q++) { Eq[e*sQ + col + QQ*q] = out[q]; }
   }
   __syncthreads();
}

// y and x pencils of the transposed L2 interpolation: t2[e][lz][qy][qx] -> out (global, [e][NL])
template<int L1D, int Q1D>
__device__ __forceinline__ void l2_values_t_yx(const double *BL, int nel, const double *t2, int s2,
                                               double *t1, int s1, double *out, int tid, int nthr)
{
   constexpr int LL = L1D*L1D, NL = LL*L1D, QP = Q1D | 1;
   for (int it = tid; it < nel*L1D*Q1D; it += nthr)       // y pencils (lz, qx)
   {
      const int e = it / (L1D*Q1D), r = it - e*(L1D*Q1D);
      const int qx = r % Q1D, lz = r / Q1D;
      double in[Q1D], o[L1D];
#pragma unroll
      for (int q = 0; q < Q1D; q++) { in[q] = t2[e*s2 + qx + Q1D*(q + Q1D*lz)]; }
      pencil_bwd<L1D,Q1D>(BL, in, o);
#pragma unroll
      for (int l = 0; l < L1D; l++) { t1[e*s1 + qx + QP*(l + L1D*lz)] = o[l]; }   // [lz][ly][qx], rows padded to QP
   }
   __syncthreads();
   for (int it = tid; it < nel*LL; it += nthr)            // x pencils (lz, ly)
   {
      const int e = it / LL, r = it - e*LL;
      double in[Q1D], o[L1D];
#pragma unroll
      for (int q = 0; q < Q1D; q++) { in[q] = t1[e*s1 + q + QP*r]; }
      pencil_bwd<L1D,Q1D>(BL, in, o);
#pragma unroll
      for (int l = 0; l < L1D; l++) { out[(size_t)e*NL + l + L1D*r] = o[l]; }
   }
}

// ---------------------------------------------------------------------------
// x and y pencils of the H1 gradient for NF fields per element:
// Xs[e][f][ND] -> Bx,Gx[e][f][dz][dy][qx] -> BB,GB,BG[e][f][dz][qy][qx].  Ends with a barrier.
// ---------------------------------------------------------------------------
template<int D1D, int Q1D, int NF>
__device__ __forceinline__ void grad_xy(const double *B, const double *G, int nel, const double *Xs, int sX,
                                        double *Bx, double *Gx, int s1, double *BB, double *GB, double *BG, int s2,
                                        int tid, int nthr)
{
   // rows of Xs and Bx/Gx padded to odd strides (DP, QP): no bank conflicts with lanes along the row index
   constexpr int DD = D1D*D1D, QQ = Q1D*Q1D, QP = Q1D | 1, DP = D1D | 1;
   for (int it = tid; it < nel*NF*DD; it += nthr)         // x pencils (f, dz, dy)
   {
      const int e = it / (NF*DD), r = it - e*(NF*DD);     // r = dy + D*(dz + D*f)
      double in[D1D], b[Q1D], g[Q1D];
#pragma unroll
      for (int d = 0; d < D1D; d++) { in[d] = Xs[e*sX + d + DP*r]; }
      pencil_fwd<D1D,Q1D>(B, in, b);
      pencil_fwd<D1D,Q1D>(G, in, g);
#pragma unroll
      for (int q = 0; q < Q1D; q++) { Bx[e*s1 + q + QP*r] = b[q]; Gx[e*s1 + q + QP*r] = g[q]; }
   }
   __syncthreads();
   for (int it = tid; it < nel*NF*D1D*Q1D; it += nthr)    // y pencils (f, dz, qx)
   {
      const int e = it / (NF*D1D*Q1D), r = it - e*(NF*D1D*Q1D);
      const int qx = r % Q1D, fz = r / Q1D;               // fz = dz + D*f
      double xb[D1D], xg[D1D], bb[Q1D], gb[Q1D], bg[Q1D];
#pragma unroll
      for (int d = 0; d < D1D; d++)
      {
         xb[d] = Bx[e*s1 + qx + QP*(d + D1D*fz)];
         xg[d] = Gx[e*s1 + qx + QP*(d + D1D*fz)];
      }
      pencil_fwd<D1D,Q1D>(B, xb, bb);
      pencil_fwd<D1D,Q1D>(B, xg, gb);
      pencil_fwd<D1D,Q1D>(G, xb, bg);
#pragma unroll
      for (int q = 0; q < Q1D; q++)
      {
         const int o = e*s2 + qx + Q1D*q + QQ*fz;        // [f][dz][qy][qx]
         BB[o] = bb[q]; GB[o] = gb[q]; BG[o] = bg[q];
      }
   }
   __syncthreads();
}

// ---------------------------------------------------------------------------
// QUpdate: one element per CTA, NT threads, one (or more) quadrature points per thread
// ---------------------------------------------------------------------------
template<int D1D, int Q1D>
struct QUpd3DCfg
{
   static constexpr int L1D = D1D - 1, DD = D1D*D1D, QQ = Q1D*Q1D, ND = D1D*DD, NQ = Q1D*QQ, NL = L1D*L1D*L1D;
   static constexpr int NF = 6;
   static constexpr int QP = Q1D | 1, DP = D1D | 1, LP = L1D | 1;   // odd row strides (bank-conflict free pencils)
   static constexpr int S_ST1 = NF*DD*QP;                  // each of Bx, Gx
   static constexpr int S_ST2 = NF*D1D*QQ;                 // each of BB, GB, BG
   static constexpr int S_E1 = L1D*L1D*QP, S_E2 = L1D*QQ;
   static constexpr int S_DOF = NF*DD*DP + L1D*L1D*LP;
   // dofs alias the stage-2 arrays (dead after the x pencils)
   static constexpr int S_A = (3*S_ST2 > S_DOF) ? 3*S_ST2 : S_DOF;
   static constexpr int S_TAB = 2*Q1D*D1D + Q1D*L1D;       // B, G, BL for the per-point z pass (runtime qz)
   // the element's Jac0inv (9 NQ) and rho0DetJ0w (NQ) slabs, brought in by bulk asynchronous copies under
   // the gather and the pencil stages; both are 16-byte multiples when NQ is even
   // (up to Q1D = 8: at Q1D = 10 the 80 KB slab would leave one resident CTA per SM)
   static constexpr bool BULK = (NQ % 2 == 0) && (NQ <= 512);
   static constexpr int S_SLAB = BULK ? 10*NQ : 0;
   static constexpr int SMEM_DOUBLES = S_A + 2*S_ST1 + S_E1 + S_E2 + 32 + S_TAB + S_SLAB + 2;
   static constexpr size_t SMEM_BYTES = sizeof(double)*SMEM_DOUBLES;
};

template<int D1D, int Q1D, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
qupdate3d(const __grid_constant__ DevTables<D1D,Q1D> tab, const int NE, const int64_t ndofs,
          const int *__restrict__ map, const double *__restrict__ S,
          const double *__restrict__ rho0DetJ0w, const double *__restrict__ Jac0inv,
          const double *__restrict__ gamma, const double *__restrict__ qweights,
          const double *__restrict__ inv_qweights,
          const QPointParams prm, double *__restrict__ sJit, double *__restrict__ dt_block_min)
{
   using C = QUpd3DCfg<D1D,Q1D>;
   extern __shared__ __align__(16) double smem[];
   double *A = smem;                            // dofs, later BB | GB | BG
   double *Bx = A + C::S_A, *Gx = Bx + C::S_ST1;
   double *E1 = Gx + C::S_ST1, *E2 = E1 + C::S_E1, *red = E2 + C::S_E2;
   double *TB = red + 32, *TG = TB + Q1D*D1D, *TBL = TG + Q1D*D1D;
   // slabs at an even double offset from the (16-byte aligned) dynamic shared memory base
   constexpr int SLAB_OFF = ((C::S_A + 2*C::S_ST1 + C::S_E1 + C::S_E2 + 32 + C::S_TAB + 1)/2)*2;
   double *J0s = smem + SLAB_OFF, *Rs = J0s + 9*C::NQ;
   __shared__ uint64_t mbar;
   const int tid = threadIdx.x;
   const int e = blockIdx.x;
   const size_t NEQ = (size_t)NE*C::NQ;
   if (C::BULK)
   {
      if (tid == 0) { mbar_init(&mbar, 1); }
      __syncthreads();
      if (tid == 0)
      {
         mbar_expect_tx(&mbar, 10*C::NQ*(unsigned int)sizeof(double));
         bulk_g2s(J0s, Jac0inv + (size_t)e*C::NQ*9, 9*C::NQ*(unsigned int)sizeof(double), &mbar);
         bulk_g2s(Rs, rho0DetJ0w + (size_t)e*C::NQ, C::NQ*(unsigned int)sizeof(double), &mbar);
      }
   }
   for (int i = tid; i < Q1D*D1D; i += NT) { TB[i] = tab.B[i]; TG[i] = tab.G[i]; }
   for (int i = tid; i < Q1D*C::L1D; i += NT) { TBL[i] = tab.BL[i]; }
   // gather x, v (6 scalar fields: S = (x | v | e), field f at offset f*ndofs) and e
   {
      const double *en = S + 6*ndofs;
      const int *m = map + (size_t)e*C::ND;
      for (int it = tid; it < C::NF*C::ND; it += NT)
      {
         const int i = it % C::ND, f = it / C::ND;
         A[i % D1D + C::DP*(i / D1D + C::DD*f)] = S[(size_t)f*ndofs + __ldg(m + i)];
      }
      for (int it = tid; it < C::NL; it += NT)
      { A[C::NF*C::DD*C::DP + it % C::L1D + C::LP*(it / C::L1D)] = en[(size_t)e*C::NL + it]; }
   }
   __syncthreads();
   // Two merged pencil stages (H1 gradient of the 6 fields + L2 interpolation of e) instead of five:
   // every barrier of a one-element CTA costs the latency tail of its slowest warp.
   double *BB = A, *GB = A + C::S_ST2, *BG = A + 2*C::S_ST2;
   {
      constexpr int DD = C::DD, QQ = C::QQ, L1D = C::L1D, LL = L1D*L1D, QP = C::QP, DP = C::DP, LP = C::LP;
      const double *Es = A + C::NF*DD*DP;
      // stage alpha: x pencils
      constexpr int nGa = C::NF*DD, nLa = LL;
      for (int it = tid; it < nGa + nLa; it += NT)
      {
         if (it < nGa)
         {
            double in[D1D], bo[Q1D], go[Q1D];
#pragma unroll
            for (int d = 0; d < D1D; d++) { in[d] = A[d + DP*it]; }
            pencil_fwd<D1D,Q1D>(tab.B, in, bo);
            pencil_fwd<D1D,Q1D>(tab.G, in, go);
#pragma unroll
            for (int q = 0; q < Q1D; q++) { Bx[q + QP*it] = bo[q]; Gx[q + QP*it] = go[q]; }
         }
         else
         {
            const int r = it - nGa;                   // ly + L1D*lz
            double in[L1D], out[Q1D];
#pragma unroll
            for (int l = 0; l < L1D; l++) { in[l] = Es[l + LP*r]; }
            pencil_fwd<L1D,Q1D>(tab.BL, in, out);
#pragma unroll
            for (int q = 0; q < Q1D; q++) { E1[q + QP*r] = out[q]; }   // [lz][ly][qx]
         }
      }
      __syncthreads();
      // stage beta: y pencils (BB | GB | BG overwrite the dofs, which are dead now)
      constexpr int nGb = C::NF*D1D*Q1D, nLb = L1D*Q1D;
      for (int it = tid; it < nGb + nLb; it += NT)
      {
         if (it < nGb)
         {
            const int qx = it % Q1D, fz = it / Q1D;   // fz = dz + D1D*f
            double xb[D1D], xg[D1D], bb[Q1D], gb[Q1D], bg[Q1D];
#pragma unroll
            for (int d = 0; d < D1D; d++)
            {
               xb[d] = Bx[qx + QP*(d + D1D*fz)];
               xg[d] = Gx[qx + QP*(d + D1D*fz)];
            }
            pencil_fwd<D1D,Q1D>(tab.B, xb, bb);
            pencil_fwd<D1D,Q1D>(tab.B, xg, gb);
            pencil_fwd<D1D,Q1D>(tab.G, xb, bg);
#pragma unroll
            for (int q = 0; q < Q1D; q++)
            {
               const int o = qx + Q1D*q + QQ*fz;      // [f][dz][qy][qx]
               BB[o] = bb[q]; GB[o] = gb[q]; BG[o] = bg[q];
            }
         }
         else
         {
            const int r = it - nGb;
            const int qx = r % Q1D, lz = r / Q1D;
            double in[L1D], out[Q1D];
#pragma unroll
            for (int l = 0; l < L1D; l++) { in[l] = E1[qx + QP*(l + L1D*lz)]; }
            pencil_fwd<L1D,Q1D>(tab.BL, in, out);
#pragma unroll
            for (int q = 0; q < Q1D; q++) { E2[qx + Q1D*(q + Q1D*lz)] = out[q]; }   // [lz][qy][qx]
         }
      }
      __syncthreads();
   }
   // z pass + point physics
   const double gam = gamma[e];
   double dt_min = prm.dt_in;
   if (C::BULK) { mbar_wait(&mbar, 0); }
   for (int q = tid; q < C::NQ; q += NT)
   {
      const int col = q % C::QQ, qz = q / C::QQ;
      double bz[D1D], gz[D1D];
#pragma unroll
      for (int dz = 0; dz < D1D; dz++) { bz[dz] = TB[qz + Q1D*dz]; gz[dz] = TG[qz + Q1D*dz]; }
      double J[9], dV[9];
#pragma unroll
      for (int f = 0; f < 6; f++)
      {
         double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
         for (int dz = 0; dz < D1D; dz++)
         {
            const int o = col + C::QQ*(dz + D1D*f);
            g0 += bz[dz]*GB[o]; g1 += bz[dz]*BG[o]; g2 += gz[dz]*BB[o];
         }
         if (f < 3) { J[f] = g0; J[f + 3] = g1; J[f + 6] = g2; }
         else { dV[f - 3] = g0; dV[f] = g1; dV[f + 3] = g2; }
      }
      const size_t eq = (size_t)e*C::NQ + q;
      double sJ[9];
      const double *j0 = C::BULK ? J0s + 9*q : Jac0inv + eq*9;
      double e_q = 0.0;
#pragma unroll
      for (int lz = 0; lz < C::L1D; lz++) { e_q += TBL[qz + Q1D*lz]*E2[col + C::QQ*lz]; }
      const double dtq = qpoint<3>(J, dV, e_q, C::BULK ? Rs[q] : __ldg(rho0DetJ0w + eq), j0, gam, __ldg(qweights + q),
                                   __ldg(inv_qweights + q), prm, sJ);
      dt_min = fmin(dt_min, dtq);
      // one 64-bit address, advanced by the plane stride (the indexed form costs ~8 integer instructions per store)
      double *sp = sJit + eq;
#pragma unroll
      for (int vd = 0; vd < 3; vd++)
#pragma unroll
         for (int gd = 0; gd < 3; gd++) { *sp = sJ[vd + gd*3]; sp += NEQ; }
   }
   // block minimum (exact, order independent); NT is a whole number of warps
   for (int o = 16; o > 0; o >>= 1) { dt_min = fmin(dt_min, __shfl_xor_sync(0xffffffffu, dt_min, o)); }
   if ((tid & 31) == 0) { red[tid >> 5] = dt_min; }
   __syncthreads();
   if (tid == 0)
   {
      double m = red[0];
      for (int w = 1; w < NT/32; w++) { m = fmin(m, red[w]); }
      atomic_min_nonneg(dt_block_min, m);   // dt_block_min: the context's running dt estimate
   }
}

// ---------------------------------------------------------------------------
// Force (L2 -> H1 vector), NB elements per CTA
// ---------------------------------------------------------------------------
template<int D1D, int Q1D>
struct Force3DCfg
{
   static constexpr int L1D = D1D - 1, DD = D1D*D1D, QQ = Q1D*Q1D, ND = D1D*DD, NQ = Q1D*QQ, NL = L1D*L1D*L1D;
   static constexpr int QP = Q1D | 1, DP = D1D | 1, LP = L1D | 1;   // odd row strides (bank-conflict free pencils)
   static constexpr int S_W = 9*D1D*QQ;          // W[c][g][dz][qy][qx]; later the element result [c][DD rows of DP]
   static constexpr int S_V = 3*DD*QP;           // each of VA, VB [c][dz][dy][qx]
   static constexpr int S_ES = L1D*L1D*LP;       // the element's L2 dofs, rows padded
   static constexpr int S_E1 = L1D*L1D*QP, S_E2 = L1D*QQ;
   // region R1 = W ; region R2 = max(VA|VB, Es|E1|E2|Eq) (the L2 scratch is dead after the z pass)
   static constexpr int S_L2 = S_ES + S_E1 + S_E2 + NQ;
   static constexpr int S_R2 = (2*S_V > S_L2) ? 2*S_V : S_L2;
   static constexpr int PER_ELEM = S_W + S_R2;
};

template<int D1D, int Q1D, int NB, int NT, bool PREFETCH>
__global__ void __launch_bounds__(NT)
force3d(const __grid_constant__ DevTables<D1D,Q1D> tab, const int NE, const int64_t ndofs,
        const int *__restrict__ map, const double *__restrict__ sJit,
        const double *__restrict__ x, double *__restrict__ y)
{
   using C = Force3DCfg<D1D,Q1D>;
   extern __shared__ double smem[];
   constexpr int PE = C::PER_ELEM, QQ = C::QQ, DD = C::DD;
   double *W = smem;                             // [e][S_W]
   double *R2 = smem + C::S_W;                   // [e][S_R2], element stride PE for both
   double *Es = R2, *E1 = Es + C::S_ES, *E2 = E1 + C::S_E1, *Eq = E2 + C::S_E2;
   double *VA = R2, *VB = R2 + C::S_V;
   constexpr int QP = C::QP, DP = C::DP;
   const int tid = threadIdx.x;
   const int eb = blockIdx.x*NB;
   const int nel = min(NB, NE - eb);
   const size_t NEQ = (size_t)NE*C::NQ;
   double *Ss = smem + NB*PE;                    // [e][cg][q], PREFETCH only
   // restriction indices of this thread's scatter items: requested first, used last (ncu: 10 % of the kernel's stall
   // samples sat on this load when it was issued inside the scatter loop at the end of the CTA)
   constexpr int NSC = (NB*3*C::ND + NT - 1)/NT;
   int sidx[NSC];
#pragma unroll
   for (int k = 0; k < NSC; k++)
   {
      const int it = tid + k*NT;
      const int e = it / (3*C::ND), r = it - e*(3*C::ND);
      sidx[k] = (it < nel*3*C::ND) ? __ldg(map + (size_t)(eb + e)*C::ND + r % C::ND) : 0;
   }
   if (PREFETCH)
   {
      static_assert(C::NQ % 2 == 0 && (NB*PE) % 2 == 0, "16-byte chunks");
      constexpr int NCH = 9*C::NQ/2;
      for (int it = tid; it < nel*NCH; it += NT)
      {
         const int e = it / NCH, r = it - e*NCH;
         const int cg = r / (C::NQ/2), h = r - cg*(C::NQ/2);
         cp_async16(Ss + (size_t)e*9*C::NQ + cg*C::NQ + 2*h, sJit + (size_t)(eb + e)*C::NQ + NEQ*cg + 2*h);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
   }
   for (int it = tid; it < nel*C::NL; it += NT)
   {
      const int e = it / C::NL, i = it - e*C::NL;
      Es[e*PE + i % C::L1D + C::LP*(i / C::L1D)] = x[(size_t)(eb + e)*C::NL + i];
   }
   __syncthreads();
   l2_values<C::L1D,Q1D>(tab.BL, nel, Es, PE, E1, PE, E2, PE, Eq, PE, tid, NT);
   if (PREFETCH)
   {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
   }
   // z pencils (e, c, g, column): W[c][g][dz][col] = sum_qz Tz(qz,dz) sJit(q,g,c) Eq(q), Tz = G for g == 2
   if (NB == 1 && !PREFETCH)
   {
      // one element per CTA: every stressJinvT load of the thread is issued before the first
      // dependent instruction (all rounds unrolled): memory-level parallelism for the HBM stream
      constexpr int NR = (9*QQ + NT - 1)/NT;
      double sv[NR][Q1D];
#pragma unroll
      for (int k = 0; k < NR; k++)
      {
         const int r = tid + k*NT;
         if (r < 9*QQ)
         {
            const int col = r % QQ, cg = r / QQ;
            const double *s = sJit + (size_t)eb*C::NQ + NEQ*cg + col;
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++) { sv[k][qz] = __ldg(s + QQ*qz); }
         }
      }
#pragma unroll
      for (int k = 0; k < NR; k++)
      {
         const int r = tid + k*NT;
         if (r < 9*QQ)
         {
            const int col = r % QQ, cg = r / QQ;
            double w[D1D];
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++) { sv[k][qz] *= Eq[col + QQ*qz]; }
            if (cg % 3 == 2) { pencil_bwd<D1D,Q1D>(tab.G, sv[k], w); }
            else { pencil_bwd<D1D,Q1D>(tab.B, sv[k], w); }
#pragma unroll
            for (int dz = 0; dz < D1D; dz++) { W[col + QQ*(dz + D1D*cg)] = w[dz]; }
         }
      }
   }
   else
   for (int it = tid; it < nel*9*QQ; it += NT)
   {
      const int e = it / (9*QQ), r = it - e*(9*QQ);
      const int col = r % QQ, cg = r / QQ;      // cg = g + 3*c
      const double *s = sJit + (size_t)(eb + e)*C::NQ + NEQ*cg + col;
      const double *ss = Ss + (size_t)e*9*C::NQ + cg*C::NQ + col;
      double sv[Q1D], w[D1D];
#pragma unroll
      for (int qz = 0; qz < Q1D; qz++) { sv[qz] = PREFETCH ? ss[QQ*qz] : __ldg(s + QQ*qz); }
#pragma unroll
      for (int qz = 0; qz < Q1D; qz++) { sv[qz] *= Eq[e*PE + col + QQ*qz]; }
      if (cg % 3 == 2) { pencil_bwd<D1D,Q1D>(tab.G, sv, w); }
      else { pencil_bwd<D1D,Q1D>(tab.B, sv, w); }
#pragma unroll
      for (int dz = 0; dz < D1D; dz++) { W[e*PE + col + QQ*(dz + D1D*cg)] = w[dz]; }
   }
   __syncthreads();
   // y pencils (e, c, dz, qx): VA = By W_c0 (x pass: G), VB = Gy W_c1 + By W_c2 (x pass: B)
   for (int it = tid; it < nel*3*D1D*Q1D; it += NT)
   {
      const int e = it / (3*D1D*Q1D), r = it - e*(3*D1D*Q1D);
      const int qx = r % Q1D, dz = (r / Q1D) % D1D, c = r / (Q1D*D1D);
      double w0[Q1D], w1[Q1D], w2[Q1D], a[D1D], b1[D1D], b2[D1D];
#pragma unroll
      for (int qy = 0; qy < Q1D; qy++)
      {
         const int o = e*PE + qx + Q1D*qy + QQ*dz;
         w0[qy] = W[o + QQ*D1D*(0 + 3*c)]; w1[qy] = W[o + QQ*D1D*(1 + 3*c)]; w2[qy] = W[o + QQ*D1D*(2 + 3*c)];
      }
      pencil_bwd<D1D,Q1D>(tab.B, w0, a);
      pencil_bwd<D1D,Q1D>(tab.G, w1, b1);
      pencil_bwd<D1D,Q1D>(tab.B, w2, b2);
#pragma unroll
      for (int dy = 0; dy < D1D; dy++)
      {
         const int o = e*PE + qx + QP*(dy + D1D*(dz + D1D*c));     // [c][dz][dy][qx], rows padded to QP
         VA[o] = a[dy]; VB[o] = b1[dy] + b2[dy];
      }
   }
   __syncthreads();
   // x pencils (e, c, dz, dy): element result [c][ND] parked in W (dead)
   const double eps2 = DBL_EPSILON*DBL_EPSILON;
   for (int it = tid; it < nel*3*DD; it += NT)
   {
      const int e = it / (3*DD), r = it - e*(3*DD);       // r = dy + D*(dz + D*c)
      double va[Q1D], vb[Q1D], oa[D1D], ob[D1D];
#pragma unroll
      for (int qx = 0; qx < Q1D; qx++) { va[qx] = VA[e*PE + qx + QP*r]; vb[qx] = VB[e*PE + qx + QP*r]; }
      pencil_bwd<D1D,Q1D>(tab.G, va, oa);
      pencil_bwd<D1D,Q1D>(tab.B, vb, ob);
#pragma unroll
      for (int dx = 0; dx < D1D; dx++)
      {
         double o = oa[dx] + ob[dx];
         if (fabs(o) < eps2) { o = 0.0; }                 // reference laghos_assembly.cpp:495-512
         W[e*PE + dx + DP*r] = o;
      }
   }
   __syncthreads();
   // scatter-add, lanes along the element-local dof index
#pragma unroll
   for (int k = 0; k < NSC; k++)
   {
      const int it = tid + k*NT;
      if (it < nel*3*C::ND)
      {
         const int e = it / (3*C::ND), r = it - e*(3*C::ND);
         const int i = r % C::ND, c = r / C::ND;
         atomicAdd(y + (size_t)c*ndofs + sidx[k], W[e*PE + i % D1D + DP*(i / D1D + DD*c)]);
      }
   }
}

// ---------------------------------------------------------------------------
// Force transpose (H1 vector -> L2), NB elements per CTA
// ---------------------------------------------------------------------------
template<int D1D, int Q1D>
struct ForceT3DCfg
{
   static constexpr int L1D = D1D - 1, DD = D1D*D1D, QQ = Q1D*Q1D, ND = D1D*DD, NQ = Q1D*QQ, NL = L1D*L1D*L1D;
   static constexpr int NF = 3;
   static constexpr int QP = Q1D | 1, DP = D1D | 1;      // odd row strides (bank-conflict free pencils)
   static constexpr int S_ST1 = NF*DD*QP, S_ST2 = NF*D1D*QQ;
   static constexpr int S_E1 = L1D*L1D*QP, S_E2 = L1D*QQ;
   // region A: Vs -> BB|GB|BG -> t1 ; region B: Bx|Gx -> t2
   static constexpr int S_A = (3*S_ST2 > NF*DD*DP) ? 3*S_ST2 : NF*DD*DP;
   static constexpr int S_B = (2*S_ST1 > S_E2) ? 2*S_ST1 : S_E2;
   static constexpr int PER_ELEM = S_A + S_B;
   static constexpr int S_PF = 9*NQ;             // stressJinvT slab of one element (prefetch variant)
};

template<int D1D, int Q1D, int NB, int NT, bool PREFETCH>
__global__ void __launch_bounds__(NT)
forcet3d(const __grid_constant__ DevTables<D1D,Q1D> tab, const int NE, const int64_t ndofs,
         const int *__restrict__ map, const double *__restrict__ sJit,
         const double *__restrict__ v, double *__restrict__ eout)
{
   using C = ForceT3DCfg<D1D,Q1D>;
   extern __shared__ double smem[];
   constexpr int PE = C::PER_ELEM, QQ = C::QQ;
   double *RA = smem, *RB = smem + C::S_A;
   double *Vs = RA, *BB = RA, *GB = RA + C::S_ST2, *BG = RA + 2*C::S_ST2, *t1 = RA;
   double *Bx = RB, *Gx = RB + C::S_ST1, *t2 = RB;
   const int tid = threadIdx.x;
   const int eb = blockIdx.x*NB;
   const int nel = min(NB, NE - eb);
   const size_t NEQ = (size_t)NE*C::NQ;
   double *Ss = smem + NB*PE;                    // [e][cg][q], PREFETCH only
   if (PREFETCH)
   {
      // the element's 9 stressJinvT planes stream into shared memory while the gather and
      // the x/y pencils run (NQ is even and the planes are 16-byte aligned)
      static_assert(C::NQ % 2 == 0 && (NB*PE) % 2 == 0, "16-byte chunks");
      constexpr int NCH = 9*C::NQ/2;
      for (int it = tid; it < nel*NCH; it += NT)
      {
         const int e = it / NCH, r = it - e*NCH;
         const int cg = r / (C::NQ/2), h = r - cg*(C::NQ/2);
         cp_async16(Ss + (size_t)e*C::S_PF + cg*C::NQ + 2*h, sJit + (size_t)(eb + e)*C::NQ + NEQ*cg + 2*h);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
   }
   for (int it = tid; it < nel*C::NF*C::ND; it += NT)
   {
      const int e = it / (C::NF*C::ND), r = it - e*(C::NF*C::ND);
      const int i = r % C::ND, c = r / C::ND;
      Vs[e*PE + i % D1D + C::DP*(i / D1D + C::DD*c)] = v[(size_t)c*ndofs + __ldg(map + (size_t)(eb + e)*C::ND + i)];
   }
   __syncthreads();
   grad_xy<D1D,Q1D,C::NF>(tab.B, tab.G, nel, Vs, PE, Bx, Gx, PE, BB, GB, BG, PE, tid, NT);
   if (PREFETCH)
   {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
   }
   // z pencils (e, column): gradients at the Q1D points of the column, contraction with
   // stressJinvT, then the z pencil of the transposed L2 interpolation in registers
   for (int it = tid; it < nel*QQ; it += NT)
   {
      const int e = it / QQ, col = it - e*QQ;
      const double *s = sJit + (size_t)(eb + e)*C::NQ + col;
      double acc[Q1D];
#pragma unroll
      for (int qz = 0; qz < Q1D; qz++) { acc[qz] = 0.0; }
#pragma unroll
      for (int c = 0; c < 3; c++)
      {
         double bb[D1D], gb[D1D], bg[D1D], g0[Q1D], g1[Q1D], g2[Q1D];
#pragma unroll
         for (int dz = 0; dz < D1D; dz++)
         {
            const int o = e*PE + col + QQ*(dz + D1D*c);
            bb[dz] = BB[o]; gb[dz] = GB[o]; bg[dz] = BG[o];
         }
         pencil_fwd<D1D,Q1D>(tab.B, gb, g0);
         pencil_fwd<D1D,Q1D>(tab.B, bg, g1);
         pencil_fwd<D1D,Q1D>(tab.G, bb, g2);
#pragma unroll
         for (int qz = 0; qz < Q1D; qz++)
         {
            // same association as the reference (:889-899): per component, sum over g, then add
            if (PREFETCH)
            {
               const double *sq = Ss + (size_t)e*C::S_PF + col + QQ*qz;
               acc[qz] += g0[qz]*sq[C::NQ*(0 + 3*c)] + g1[qz]*sq[C::NQ*(1 + 3*c)] + g2[qz]*sq[C::NQ*(2 + 3*c)];
            }
            else
            {
               const double *sq = s + QQ*qz;
               acc[qz] += g0[qz]*__ldg(sq + NEQ*(0 + 3*c)) + g1[qz]*__ldg(sq + NEQ*(1 + 3*c)) + g2[qz]*__ldg(sq + NEQ*(2 + 3*c));
            }
         }
      }
      double o[C::L1D];
      pencil_bwd<C::L1D,Q1D>(tab.BL, acc, o);
#pragma unroll
      for (int lz = 0; lz < C::L1D; lz++) { t2[e*PE + col + QQ*lz] = o[lz]; }   // Bx|Gx are dead
   }
   __syncthreads();
   l2_values_t_yx<C::L1D,Q1D>(tab.BL, nel, t2, PE, t1, PE, eout + (size_t)eb*C::NL, tid, NT);
}

// ---------------------------------------------------------------------------
// Force transpose, persistent: a CTA walks over element batches bi = blockIdx.x, + gridDim.x, ... and keeps the
// gather of the NEXT batch in flight (values in registers, the indices of the batch after that as well) while it
// works on the current one.  ncu of the one-batch-per-CTA kernel: 27 % of all stall samples sit on the dependent
// index -> value loads of the gather at the start of every 9-us CTA, another 18 % on the barriers behind it.
// The stressJinvT slab of the batch is bulk-prefetched (cp.async) as before.  Measured 1358 vs 1452 us (cfg 2).
// Tried on top and dropped: z pencils split over the three components with a 3-lane shuffle reduction (108 busy
// threads instead of 36 in that stage): 1607 us with 128 threads, 2039 us with 96.
// ---------------------------------------------------------------------------
template<int D1D, int Q1D, int NB, int NT>
__global__ void __launch_bounds__(NT)
forcet3d_persist(const __grid_constant__ DevTables<D1D,Q1D> tab, const int NE, const int64_t ndofs,
                 const int *__restrict__ map, const double *__restrict__ sJit,
                 const double *__restrict__ v, double *__restrict__ eout)
{
   using C = ForceT3DCfg<D1D,Q1D>;
   extern __shared__ double smem[];
   constexpr int PE = C::PER_ELEM, QQ = C::QQ;
   double *RA = smem, *RB = smem + C::S_A;
   double *Vs = RA, *BB = RA, *GB = RA + C::S_ST2, *BG = RA + 2*C::S_ST2, *t1 = RA;
   double *Bx = RB, *Gx = RB + C::S_ST1, *t2 = RB;
   double *Ss = smem + NB*PE;                    // [e][cg][q]
   static_assert(C::NQ % 2 == 0 && (NB*PE) % 2 == 0, "16-byte chunks");
   const int tid = threadIdx.x;
   const size_t NEQ = (size_t)NE*C::NQ;
   const int nbatch = (NE + NB - 1)/NB;
   constexpr int PER = NB*C::NF*C::ND;            // gathered values per batch
   constexpr int NIT = (PER + NT - 1)/NT;
   int idx_n[NIT];                                // restriction indices of the batch after the next
   double xr[NIT];                                // gathered values of the next batch
   auto load_idx = [&](int bi)
   {
      const int nd = min(NB, NE - bi*NB)*C::NF*C::ND;
#pragma unroll
      for (int k = 0; k < NIT; k++)
      {
         const int it = tid + k*NT;
         const int e = it / (C::NF*C::ND), r = it - e*(C::NF*C::ND);
         idx_n[k] = (bi < nbatch && it < nd) ? __ldg(map + (size_t)(bi*NB + e)*C::ND + r % C::ND) : 0;
      }
   };
   auto load_val = [&](int bi)                    // uses idx_n (loaded one batch earlier)
   {
      const int nd = min(NB, NE - bi*NB)*C::NF*C::ND;
#pragma unroll
      for (int k = 0; k < NIT; k++)
      {
         const int it = tid + k*NT;
         const int r = it % (C::NF*C::ND);
         xr[k] = (bi < nbatch && it < nd) ? v[(size_t)(r / C::ND)*ndofs + idx_n[k]] : 0.0;
      }
   };
   int bi = blockIdx.x;
   load_idx(bi);
   load_val(bi);
   load_idx(bi + gridDim.x);
   for (; bi < nbatch; bi += gridDim.x)
   {
      const int eb = bi*NB;
      const int nel = min(NB, NE - eb);
      {
         constexpr int NCH = 9*C::NQ/2;
         for (int it = tid; it < nel*NCH; it += NT)
         {
            const int e = it / NCH, r = it - e*NCH;
            const int cg = r / (C::NQ/2), h = r - cg*(C::NQ/2);
            cp_async16(Ss + (size_t)e*C::S_PF + cg*C::NQ + 2*h, sJit + (size_t)(eb + e)*C::NQ + NEQ*cg + 2*h);
         }
         asm volatile("cp.async.commit_group;" ::: "memory");
      }
#pragma unroll
      for (int k = 0; k < NIT; k++)
      {
         const int it = tid + k*NT;
         if (it < nel*C::NF*C::ND)
         {
            const int e = it / (C::NF*C::ND), r = it - e*(C::NF*C::ND);
            const int i = r % C::ND, c = r / C::ND;
            Vs[e*PE + i % D1D + C::DP*(i / D1D + C::DD*c)] = xr[k];
         }
      }
      __syncthreads();
      // next batch: values through the indices that arrived during the previous batch, then the indices after that
      load_val(bi + gridDim.x);
      load_idx(bi + 2*gridDim.x);
      grad_xy<D1D,Q1D,C::NF>(tab.B, tab.G, nel, Vs, PE, Bx, Gx, PE, BB, GB, BG, PE, tid, NT);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      for (int it = tid; it < nel*QQ; it += NT)
      {
         const int e = it / QQ, col = it - e*QQ;
         double acc[Q1D];
#pragma unroll
         for (int qz = 0; qz < Q1D; qz++) { acc[qz] = 0.0; }
#pragma unroll
         for (int c = 0; c < 3; c++)
         {
            double bb[D1D], gb[D1D], bg[D1D], g0[Q1D], g1[Q1D], g2[Q1D];
#pragma unroll
            for (int dz = 0; dz < D1D; dz++)
            {
               const int o = e*PE + col + QQ*(dz + D1D*c);
               bb[dz] = BB[o]; gb[dz] = GB[o]; bg[dz] = BG[o];
            }
            pencil_fwd<D1D,Q1D>(tab.B, gb, g0);
            pencil_fwd<D1D,Q1D>(tab.B, bg, g1);
            pencil_fwd<D1D,Q1D>(tab.G, bb, g2);
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++)
            {
               // same association as the reference (:889-899): per component, sum over g, then add
               const double *sq = Ss + (size_t)e*C::S_PF + col + QQ*qz;
               acc[qz] += g0[qz]*sq[C::NQ*(0 + 3*c)] + g1[qz]*sq[C::NQ*(1 + 3*c)] + g2[qz]*sq[C::NQ*(2 + 3*c)];
            }
         }
         double o[C::L1D];
         pencil_bwd<C::L1D,Q1D>(tab.BL, acc, o);
#pragma unroll
         for (int lz = 0; lz < C::L1D; lz++) { t2[e*PE + col + QQ*lz] = o[lz]; }   // Bx|Gx are dead
      }
      __syncthreads();
      l2_values_t_yx<C::L1D,Q1D>(tab.BL, nel, t2, PE, t1, PE, eout + (size_t)eb*C::NL, tid, NT);
      __syncthreads();     // RA / RB / Ss are rewritten by the next batch
   }
}

// ---------------------------------------------------------------------------
// L2 mass apply: y_e = BL^t D BL x_e (block diagonal), NB elements per CTA
// ---------------------------------------------------------------------------
template<int D1D, int Q1D>
struct MassL2Cfg
{
   static constexpr int L1D = D1D - 1, QQ = Q1D*Q1D, NQ = Q1D*QQ, NL = L1D*L1D*L1D;
   static constexpr int QP = Q1D | 1, LP = L1D | 1;      // odd row strides (bank-conflict free pencils)
   static constexpr int S_ES = L1D*L1D*LP;
   static constexpr int S_E1 = L1D*L1D*QP, S_E2 = L1D*QQ;
   static constexpr int PER_ELEM = S_ES + S_E1 + S_E2;
};

template<int D1D, int Q1D, int NB, int NT>
__global__ void __launch_bounds__(NT)
massl2_3d(const __grid_constant__ DevTables<D1D,Q1D> tab, const int NE,
          const double *__restrict__ Dq, const double *__restrict__ x, double *__restrict__ y)
{
   using C = MassL2Cfg<D1D,Q1D>;
   extern __shared__ double smem[];
   constexpr int PE = C::PER_ELEM, QQ = C::QQ, L1D = C::L1D, LL = L1D*L1D, QP = C::QP, LP = C::LP;
   double *Es = smem, *t1 = Es + C::S_ES, *t2 = t1 + C::S_E1;
   const int tid = threadIdx.x;
   const int eb = blockIdx.x*NB;
   const int nel = min(NB, NE - eb);
   for (int it = tid; it < nel*C::NL; it += NT)
   {
      const int e = it / C::NL, i = it - e*C::NL;
      Es[e*PE + i % L1D + LP*(i / L1D)] = x[(size_t)(eb + e)*C::NL + i];
   }
   __syncthreads();
   for (int it = tid; it < nel*LL; it += NT)              // x pencils
   {
      const int e = it / LL, r = it - e*LL;
      double in[L1D], out[Q1D];
#pragma unroll
      for (int l = 0; l < L1D; l++) { in[l] = Es[e*PE + l + LP*r]; }
      pencil_fwd<L1D,Q1D>(tab.BL, in, out);
#pragma unroll
      for (int q = 0; q < Q1D; q++) { t1[e*PE + q + QP*r] = out[q]; }
   }
   __syncthreads();
   for (int it = tid; it < nel*L1D*Q1D; it += NT)         // y pencils
   {
      const int e = it / (L1D*Q1D), r = it - e*(L1D*Q1D);
      const int qx = r % Q1D, lz = r / Q1D;
      double in[L1D], out[Q1D];
#pragma unroll
      for (int l = 0; l < L1D; l++) { in[l] = t1[e*PE + qx + QP*(l + L1D*lz)]; }
      pencil_fwd<L1D,Q1D>(tab.BL, in, out);
#pragma unroll
      for (int q = 0; q < Q1D; q++) { t2[e*PE + qx + Q1D*(q + Q1D*lz)] = out[q]; }
   }
   __syncthreads();
   for (int it = tid; it < nel*QQ; it += NT)              // z pencils: forward, scale by D, back (in place)
   {
      const int e = it / QQ, col = it - e*QQ;
      const double *d = Dq + (size_t)(eb + e)*C::NQ + col;
      double in[L1D], w[Q1D], u[Q1D], o[L1D];
#pragma unroll
      for (int q = 0; q < Q1D; q++) { w[q] = __ldg(d + QQ*q); }   // issue the loads first
#pragma unroll
      for (int l = 0; l < L1D; l++) { in[l] = t2[e*PE + col + QQ*l]; }
      pencil_fwd<L1D,Q1D>(tab.BL, in, u);
#pragma unroll
      for (int q = 0; q < Q1D; q++) { u[q] *= w[q]; }
      pencil_bwd<L1D,Q1D>(tab.BL, u, o);
#pragma unroll
      for (int l = 0; l < L1D; l++) { t2[e*PE + col + QQ*l] = o[l]; }
   }
   __syncthreads();
   l2_values_t_yx<L1D,Q1D>(tab.BL, nel, t2, PE, t1, PE, y + (size_t)eb*C::NL, tid, NT);
}

} // namespace tuned
} // namespace lagb
