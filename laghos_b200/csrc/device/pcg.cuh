// Device-resident (P)CG: MFEM CGSolver::Mult restated (call
// sites reference laghos_solver.cpp:388 and :481) with all scalars (nom, den,
// alpha, beta, r0, convergence flags) kept in device memory so that an iteration is
// a fixed sequence of kernel launches with no host round trip; NC independent
// systems (the `dim` velocity components) advance in lock-step and share one pass
// over the operator's quadrature data.
//
// Essential dofs: the Jacobi inverse diagonal `dinvm` is zero at the component's
// essential dofs, which keeps d, and therefore x, exactly zero there; entries of r
// and z at those dofs are never used (they are multiplied by zero), which is
// arithmetically identical to zeroing the operator rows and the RHS entries
// (MassPAOperator::Mult / EliminateRHS, reference laghos_assembly.cpp:112-121).
//
// The inner products are reduced in a fixed order (per CTA, then per finish-kernel chunk, then across ranks in
// ascending rank order): for given input vectors they are bit-reproducible.  The operator's E -> L scatter is not:
// it adds with red.global.add.f64, so A d (and from there the iterates) can differ in the last bits from run to run;
// iteration counts are stable to +-1 and |e| to ~1e-12 (tests/test_gpu_end_to_end.py).  The deterministic scatter
// exists (coloured brick schedule, lagb_tune_set key 6) but is slower.
#pragma once
#include "common.cuh"
#include "p2p_prims.cuh"

namespace lagb {
namespace pcg {

constexpr int MAXC = 3;
constexpr int RB = 256;       // reduction / vector kernel block size

struct State                  // lives in device memory
{
   double nom[MAXC], den[MAXC], betanom[MAXC], r0[MAXC], alpha[MAXC], beta[MAXC];
   int done[MAXC];            // 1 once the component has stopped iterating
   int iters[MAXC];           // MFEM final_iter
   int all_done;
};

// Jacobi preconditioner with the essential dofs of component `comp` eliminated: one shared
// inverse diagonal for all components plus a per-dof bit mask (bit c = essential for
// component c) instead of `dim` masked copies of the diagonal (saves 2 vector reads per pass).
struct Prec
{
   const double *dinv;            // [n] or nullptr (unpreconditioned L2 CG)
   const unsigned char *ess;      // [n] or nullptr
   int comp0;                     // component of the first system
};
__device__ __forceinline__ double prec_apply(const Prec &P, int64_t i, int c, double r)
{
   if (!P.dinv) { return r; }
   const double z = P.dinv[i]*r;
   return (P.ess && ((P.ess[i] >> (P.comp0 + c)) & 1)) ? 0.0 : z;
}

__device__ __forceinline__ double block_sum(double v, double *sh)
{
   for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
   const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
   __syncthreads();
   if ((threadIdx.x & 31) == 0) { sh[w] = v; }
   __syncthreads();
   double s = 0.0;
   if (threadIdx.x == 0) { for (int i = 0; i < nw; i++) { s += sh[i]; } }
   return s;   // valid on thread 0
}

// r = b - z (z = A x), zz = dinvm*r, d = zz, nom_part = sum own*zz*r ; z = 0
template<int NC>
__global__ void init_residual(int64_t n, int64_t cstride, const double *__restrict__ b, double *__restrict__ z,
                              const Prec P, const unsigned char *__restrict__ own,
                              double *__restrict__ r, double *__restrict__ d, double *__restrict__ part,
                              int iterative_mode)
{
   pdl_launch(); pdl_wait();
   __shared__ double sh[32];
   double acc[NC];
#pragma unroll
   for (int c = 0; c < NC; c++) { acc[c] = 0.0; }
   for (int64_t i = blockIdx.x*(int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x)
   {
      const double w = own ? (double)own[i] : 1.0;
#pragma unroll
      for (int c = 0; c < NC; c++)
      {
         const int64_t k = i + c*cstride;
         const double rr = (iterative_mode & 1) ? b[k] - z[k] : b[k];
         const double zz = prec_apply(P, i, c, rr);
         r[k] = rr; d[k] = zz;
         if (iterative_mode & 2) { z[k] = 0.0; }      // bit 1: leave a zeroed result vector behind
         acc[c] += w*zz*rr;
      }
   }
#pragma unroll
   for (int c = 0; c < NC; c++)
   {
      const double s = block_sum(acc[c], sh);
      if (threadIdx.x == 0) { part[(size_t)blockIdx.x*NC + c] = s; }
   }
}

// generic partial dot: part[b*NC + c] = sum own*x*y
template<int NC>
__global__ void dot_partial(int64_t n, int64_t cstride, const double *__restrict__ x, const double *__restrict__ y,
                            const unsigned char *__restrict__ own, double *__restrict__ part)
{
   pdl_launch(); pdl_wait();
   __shared__ double sh[32];
   double acc[NC];
#pragma unroll
   for (int c = 0; c < NC; c++) { acc[c] = 0.0; }
   for (int64_t i = blockIdx.x*(int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x)
   {
      const double w = own ? (double)own[i] : 1.0;
#pragma unroll
      for (int c = 0; c < NC; c++) { acc[c] += w*x[i + c*cstride]*y[i + c*cstride]; }
   }
#pragma unroll
   for (int c = 0; c < NC; c++)
   {
      const double s = block_sum(acc[c], sh);
      if (threadIdx.x == 0) { part[(size_t)blockIdx.x*NC + c] = s; }
   }
}

// out[c] = sum_b part[b*NC + c], fixed order; one block.
template<int NC>
__global__ void reduce_partials(int nblocks, const double *__restrict__ part, double *__restrict__ out)
{
   pdl_launch(); pdl_wait();
   __shared__ double sh[32];
#pragma unroll
   for (int c = 0; c < NC; c++)
   {
      double a = 0.0;
      for (int b = threadIdx.x; b < nblocks; b += blockDim.x) { a += part[(size_t)b*NC + c]; }
      const double s = block_sum(a, sh);
      if (threadIdx.x == 0) { out[c] = s; }
   }
}

// Fixed-order reduction of part[b*NC + c], b < nblocks, by one block of FB threads: thread t
// sums b = t, t+FB, ... with 8 independent partial sums per component and all components in one
// sweep (up to 24 loads in flight per thread: the kernel is pure load latency), then a
// shuffle/shared tree.  Result valid on thread 0.
constexpr int FB = 1024;
template<int NC>
__device__ __forceinline__ void reduce_to_thread0(const double *__restrict__ part, int nblocks, double *out, double *sh)
{
   constexpr int W = 8;
   double a[NC][W];
#pragma unroll
   for (int c = 0; c < NC; c++)
#pragma unroll
      for (int w = 0; w < W; w++) { a[c][w] = 0.0; }
   int b = threadIdx.x;
   for (; b + (W - 1)*FB < nblocks; b += W*FB)
   {
#pragma unroll
      for (int w = 0; w < W; w++)
#pragma unroll
         for (int c = 0; c < NC; c++) { a[c][w] += part[(size_t)(b + w*FB)*NC + c]; }
   }
   for (; b < nblocks; b += FB)
   {
#pragma unroll
      for (int c = 0; c < NC; c++) { a[c][0] += part[(size_t)b*NC + c]; }
   }
#pragma unroll
   for (int c = 0; c < NC; c++)
   {
      out[c] = block_sum(((a[c][0] + a[c][1]) + (a[c][2] + a[c][3])) + ((a[c][4] + a[c][5]) + (a[c][6] + a[c][7])), sh);
   }
}

// The same over a small grid: CTA b reduces the b-th contiguous chunk (fixed order), the last CTA to arrive
// (atomic ticket) adds the chunk sums in CTA order.  A single CTA needs ~17 us for the 98304 partials of the
// mass kernel at 64^3 elements (pure load latency); 32 CTAs need ~4 us.  Returns true on the CTA that holds the
// total (valid on its thread 0); deterministic.
constexpr int FIN_CTAS = 32;
template<int NC>
__device__ __forceinline__ bool reduce_grid(const double *__restrict__ part, int nblocks, double *stage, unsigned int *ctr,
                                            double *out, double *sh)
{
   if (gridDim.x == 1) { reduce_to_thread0<NC>(part, nblocks, out, sh); return true; }
   const int per = (nblocks + gridDim.x - 1)/gridDim.x;
   const int b0 = min(nblocks, (int)blockIdx.x*per), b1 = min(nblocks, b0 + per);
   reduce_to_thread0<NC>(part + (size_t)b0*NC, b1 - b0, out, sh);
   __shared__ bool last;
   __shared__ double chunk[FIN_CTAS*MAXC];
   if (threadIdx.x == 0)
   {
      for (int c = 0; c < NC; c++) { stage[blockIdx.x*NC + c] = out[c]; }
      __threadfence();
      last = (atomicAdd(ctr, 1u) == gridDim.x - 1);
   }
   __syncthreads();
   if (!last) { return false; }
   __threadfence();
   if (threadIdx.x < gridDim.x*NC) { chunk[threadIdx.x] = __ldcg(stage + threadIdx.x); }
   __syncthreads();
   if (threadIdx.x == 0)
   {
      for (int c = 0; c < NC; c++)
      {
         double s = 0.0;
         for (unsigned int b = 0; b < gridDim.x; b++) { s += chunk[b*NC + c]; }
         out[c] = s;
      }
      *ctr = 0u;
   }
   return true;
}

// multi-rank path: partials -> NC sums (input of the NCCL all-reduce)
template<int NC>
__global__ void __launch_bounds__(FB) reduce_final(int nblocks, const double *__restrict__ part, double *__restrict__ out)
{
   pdl_launch(); pdl_wait();
   __shared__ double sh[32];
   double tmp[NC];
   reduce_to_thread0<NC>(part, nblocks, tmp, sh);
   if (threadIdx.x == 0) { for (int c = 0; c < NC; c++) { out[c] = tmp[c]; } }
}

// The finish kernels take the per-block partials straight from the producing kernel
// (nblocks > 1, single rank) or the already reduced and all-reduced sums (nblocks == 1).
template<int NC>
__global__ void __launch_bounds__(FB) finish_init(State *st, const double *part, int nblocks, double rel_tol, double abs_tol,
                            const p2p::Dev *pd, unsigned long long seq, double *stage, unsigned int *ctr)
{
   pdl_launch(); pdl_wait();
   __shared__ double sh[32];
   __shared__ double shx[MAXC];
   double tmp[NC];
   if (!reduce_grid<NC>(part, nblocks, stage, ctr, tmp, sh)) { return; }
   if (pd) { p2p::allreduce_cta<NC>(*pd, seq, tmp, shx); }
   if (threadIdx.x != 0) { return; }
   int all = 1;
   for (int c = 0; c < NC; c++)
   {
      const double nom = tmp[c];
      st->nom[c] = nom;
      st->iters[c] = 0;
      st->done[c] = 0;
      st->beta[c] = 0.0; st->alpha[c] = 0.0;
      if (nom < 0.0) { st->done[c] = 1; }   // not positive definite: MFEM returns, final_iter = 0
      const double r0 = fmax(nom*rel_tol*rel_tol, abs_tol*abs_tol);
      st->r0[c] = r0;
      if (nom <= r0) { st->done[c] = 1; }
      all &= st->done[c];
   }
   st->all_done = all;
}

// den reduced into tmp: alpha = nom/den
template<int NC>
__global__ void __launch_bounds__(FB) finish_den(State *st, const double *part, int nblocks, int iter,
                           const p2p::Dev *pd, unsigned long long seq, double *stage, unsigned int *ctr)
{
   pdl_launch(); pdl_wait();
   __shared__ double sh[32];
   __shared__ double shx[MAXC];
   double tmp[NC];
   if (!reduce_grid<NC>(part, nblocks, stage, ctr, tmp, sh)) { return; }
   if (pd) { p2p::allreduce_cta<NC>(*pd, seq, tmp, shx); }
   if (threadIdx.x != 0) { return; }
   for (int c = 0; c < NC; c++)
   {
      if (st->done[c]) { st->alpha[c] = 0.0; continue; }   // stopped earlier: the x update below runs with alpha = 0
      const double den = tmp[c];
      st->den[c] = den;
      if (den == 0.0)
      {
         // MFEM: stop, final_iter = current iteration index (0 before the loop)
         st->done[c] = 1; st->iters[c] = iter - 1; st->alpha[c] = 0.0;
         continue;
      }
      st->alpha[c] = st->nom[c]/den;
   }
}

// x += alpha d ; r -= alpha z ; betanom_part = sum own*r*(M^-1 r)
// Branch-free and load-first: a finished component runs with alpha = 0 (x and r are
// rewritten unchanged), all loads of UNR grid-stride iterations are issued before the first
// dependent instruction (memory-level parallelism: the kernel is a pure HBM stream).
template<int NC>
__global__ void __launch_bounds__(RB)
update_xr(int64_t n, int64_t cstride, const State *__restrict__ st,
          double *__restrict__ x, double *__restrict__ r, const double *__restrict__ d,
          const double *__restrict__ z, const Prec P,
          const unsigned char *__restrict__ own, double *__restrict__ part)
{
   pdl_launch(); pdl_wait();
   constexpr int UNR = 2;
   __shared__ double sh[32];
   double acc[NC], alpha[NC];
#pragma unroll
   for (int c = 0; c < NC; c++) { acc[c] = 0.0; alpha[c] = st->done[c] ? 0.0 : st->alpha[c]; }
   const int64_t stride = (int64_t)gridDim.x*blockDim.x;
   for (int64_t i0 = blockIdx.x*(int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += UNR*stride)
   {
      double xv[UNR][NC], dv[UNR][NC], rv[UNR][NC], zv[UNR][NC], di[UNR], w[UNR];
      unsigned int em[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++)
      {
         const int64_t i = i0 + u*stride;
         const bool ok = i < n;
         di[u] = (ok && P.dinv) ? P.dinv[i] : 1.0;
         em[u] = (ok && P.ess) ? (unsigned int)P.ess[i] >> P.comp0 : 0u;
         w[u] = (ok && own) ? (double)own[i] : 1.0;
#pragma unroll
         for (int c = 0; c < NC; c++)
         {
            const int64_t k = i + c*cstride;
            xv[u][c] = ok ? x[k] : 0.0; dv[u][c] = ok ? d[k] : 0.0;
            rv[u][c] = ok ? r[k] : 0.0; zv[u][c] = ok ? z[k] : 0.0;
         }
      }
#pragma unroll
      for (int u = 0; u < UNR; u++)
      {
         const int64_t i = i0 + u*stride;
         if (i < n)
         {
#pragma unroll
            for (int c = 0; c < NC; c++)
            {
               const int64_t k = i + c*cstride;
               x[k] = xv[u][c] + alpha[c]*dv[u][c];
               const double rr = rv[u][c] - alpha[c]*zv[u][c];
               r[k] = rr;
               const double zz = ((em[u] >> c) & 1u) ? 0.0 : di[u]*rr;
               acc[c] += w[u]*rr*zz;
            }
         }
      }
   }
#pragma unroll
   for (int c = 0; c < NC; c++)
   {
      const double s = block_sum(acc[c], sh);
      if (threadIdx.x == 0) { part[(size_t)blockIdx.x*NC + c] = s; }
   }
}

// betanom reduced into tmp: convergence test, beta, nom <- betanom
template<int NC>
__global__ void __launch_bounds__(FB) finish_beta(State *st, const double *part, int nblocks, int iter, int max_iter,
                            const p2p::Dev *pd, unsigned long long seq, double *stage, unsigned int *ctr)
{
   pdl_launch(); pdl_wait();
   __shared__ double sh[32];
   __shared__ double shx[MAXC];
   double tmp[NC];
   if (!reduce_grid<NC>(part, nblocks, stage, ctr, tmp, sh)) { return; }
   if (pd) { p2p::allreduce_cta<NC>(*pd, seq, tmp, shx); }
   if (threadIdx.x != 0) { return; }
   int all = 1;
   for (int c = 0; c < NC; c++)
   {
      if (!st->done[c])
      {
         const double betanom = tmp[c];
         st->betanom[c] = betanom;
         if (betanom < 0.0 || betanom <= st->r0[c]) { st->done[c] = 1; st->iters[c] = iter; }
         else if (iter >= max_iter) { st->done[c] = 1; st->iters[c] = max_iter; }
         else
         {
            st->beta[c] = betanom/st->nom[c];
            st->nom[c] = betanom;
         }
      }
      all &= st->done[c];
   }
   st->all_done = all;
}

// d = M^-1 r + beta d ; z = 0   (d of a finished component is left as it is)
template<int NC>
__global__ void __launch_bounds__(RB)
update_d(int64_t n, int64_t cstride, const State *__restrict__ st,
         double *__restrict__ d, const double *__restrict__ r,
         const Prec P, double *__restrict__ z)
{
   pdl_launch(); pdl_wait();
   constexpr int UNR = 2;
   double beta[NC]; bool skip[NC];
#pragma unroll
   for (int c = 0; c < NC; c++) { beta[c] = st->beta[c]; skip[c] = st->done[c] != 0; }
   const int64_t stride = (int64_t)gridDim.x*blockDim.x;
   for (int64_t i0 = blockIdx.x*(int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += UNR*stride)
   {
      double dv[UNR][NC], rv[UNR][NC], di[UNR];
      unsigned int em[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++)
      {
         const int64_t i = i0 + u*stride;
         const bool ok = i < n;
         di[u] = (ok && P.dinv) ? P.dinv[i] : 1.0;
         em[u] = (ok && P.ess) ? (unsigned int)P.ess[i] >> P.comp0 : 0u;
#pragma unroll
         for (int c = 0; c < NC; c++)
         {
            const int64_t k = i + c*cstride;
            dv[u][c] = ok ? d[k] : 0.0; rv[u][c] = ok ? r[k] : 0.0;
         }
      }
#pragma unroll
      for (int u = 0; u < UNR; u++)
      {
         const int64_t i = i0 + u*stride;
         if (i < n)
         {
#pragma unroll
            for (int c = 0; c < NC; c++)
            {
               const int64_t k = i + c*cstride;
               const double zz = ((em[u] >> c) & 1u) ? 0.0 : di[u]*rv[u][c];
               d[k] = skip[c] ? dv[u][c] : zz + beta[c]*dv[u][c];
               z[k] = 0.0;
            }
         }
      }
   }
}

// Two-kernel split of the same arithmetic with one vector pass less per iteration: the x update
// moves from the residual kernel to the direction kernel, which reads d anyway.
//   update_r : r -= alpha z ; betanom_part = sum own*r*(M^-1 r)            (reads r, z; writes r)
//   update_dx: x += alpha d ; d = M^-1 r + beta d ; [z = 0]                (reads x, d, r; writes x, d)
// alpha of a component that stopped in an EARLIER iteration is 0 (finish_den), a component that
// converged in THIS iteration still gets its x update (MFEM updates x before the convergence test).
template<int NC>
__global__ void __launch_bounds__(RB)
update_r(int64_t n, int64_t cstride, const State *__restrict__ st,
         double *__restrict__ r, const double *__restrict__ z, const Prec P,
         const unsigned char *__restrict__ own, double *__restrict__ part)
{
   pdl_launch(); pdl_wait();
   constexpr int UNR = 2;   // 2 iterations in flight: 80 registers, 3 CTAs/SM; 4 (114 registers, 2 CTAs/SM) measured 2.4 % slower on T_cgH1
   __shared__ double sh[32];
   double acc[NC], alpha[NC];
#pragma unroll
   for (int c = 0; c < NC; c++) { acc[c] = 0.0; alpha[c] = st->done[c] ? 0.0 : st->alpha[c]; }
   const int64_t stride = (int64_t)gridDim.x*blockDim.x;
   for (int64_t i0 = blockIdx.x*(int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += UNR*stride)
   {
      double rv[UNR][NC], zv[UNR][NC], di[UNR], w[UNR];
      unsigned int em[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++)
      {
         const int64_t i = i0 + u*stride;
         const bool ok = i < n;
         di[u] = (ok && P.dinv) ? P.dinv[i] : 1.0;
         em[u] = (ok && P.ess) ? (unsigned int)P.ess[i] >> P.comp0 : 0u;
         w[u] = (ok && own) ? (double)own[i] : 1.0;
#pragma unroll
         for (int c = 0; c < NC; c++)
         {
            const int64_t k = i + c*cstride;
            rv[u][c] = ok ? r[k] : 0.0; zv[u][c] = ok ? z[k] : 0.0;
         }
      }
#pragma unroll
      for (int u = 0; u < UNR; u++)
      {
         const int64_t i = i0 + u*stride;
         if (i < n)
         {
#pragma unroll
            for (int c = 0; c < NC; c++)
            {
               const double rr = rv[u][c] - alpha[c]*zv[u][c];
               r[i + c*cstride] = rr;
               const double zz = ((em[u] >> c) & 1u) ? 0.0 : di[u]*rr;
               acc[c] += w[u]*rr*zz;
            }
         }
      }
   }
#pragma unroll
   for (int c = 0; c < NC; c++)
   {
      const double s = block_sum(acc[c], sh);
      if (threadIdx.x == 0) { part[(size_t)blockIdx.x*NC + c] = s; }
   }
}

template<int NC, bool ZERO_Z>
__global__ void __launch_bounds__(RB)
update_dx(int64_t n, int64_t cstride, const State *__restrict__ st,
          double *__restrict__ x, double *__restrict__ d, const double *__restrict__ r,
          const Prec P, double *__restrict__ z)
{
   pdl_launch(); pdl_wait();
   constexpr int UNR = 2;   // 1 iteration in flight measured 1 % slower
   double alpha[NC], beta[NC]; bool skip[NC];
#pragma unroll
   for (int c = 0; c < NC; c++) { alpha[c] = st->alpha[c]; beta[c] = st->beta[c]; skip[c] = st->done[c] != 0; }
   const int64_t stride = (int64_t)gridDim.x*blockDim.x;
   for (int64_t i0 = blockIdx.x*(int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += UNR*stride)
   {
      double xv[UNR][NC], dv[UNR][NC], rv[UNR][NC], di[UNR];
      unsigned int em[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++)
      {
         const int64_t i = i0 + u*stride;
         const bool ok = i < n;
         di[u] = (ok && P.dinv) ? P.dinv[i] : 1.0;
         em[u] = (ok && P.ess) ? (unsigned int)P.ess[i] >> P.comp0 : 0u;
#pragma unroll
         for (int c = 0; c < NC; c++)
         {
            const int64_t k = i + c*cstride;
            xv[u][c] = ok ? x[k] : 0.0; dv[u][c] = ok ? d[k] : 0.0; rv[u][c] = ok ? r[k] : 0.0;
         }
      }
#pragma unroll
      for (int u = 0; u < UNR; u++)
      {
         const int64_t i = i0 + u*stride;
         if (i < n)
         {
#pragma unroll
            for (int c = 0; c < NC; c++)
            {
               const int64_t k = i + c*cstride;
               x[k] = xv[u][c] + alpha[c]*dv[u][c];
               const double zz = ((em[u] >> c) & 1u) ? 0.0 : di[u]*rv[u][c];
               d[k] = skip[c] ? dv[u][c] : zz + beta[c]*dv[u][c];
               if (ZERO_Z) { z[k] = 0.0; }
            }
         }
      }
   }
}

// ---- plain vector kernels (shim Vector ops, RK stage combinations) ----
static __global__ void vec_fill(double *__restrict__ y, double a, int64_t n)
{
   for (int64_t i = blockIdx.x*(int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x) { y[i] = a; }
}
static __global__ void vec_axpby(double *__restrict__ z, double a, const double *__restrict__ x, double b,
                          const double *__restrict__ y, int64_t n)
{
   for (int64_t i = blockIdx.x*(int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x)
   {
      z[i] = (b == 0.0) ? a*x[i] : a*x[i] + b*y[i];
   }
}
static __global__ void vec_zero_idx(double *__restrict__ y, const int *__restrict__ idx, int n)
{
   for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < n; i += gridDim.x*blockDim.x) { y[idx[i]] = 0.0; }
}
} // namespace pcg
} // namespace lagb
