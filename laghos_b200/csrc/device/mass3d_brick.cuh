// "Brick" H1 velocity-mass partial-assembly apply:  y = G^t B^t D B G x   (no atomics)
// (reference MassPAOperator::Mult -> MFEM MassIntegrator::AddMultPA,
//  laghos_assembly.cpp:117-121; arithmetic as in amr/laghos_assembly.cpp:878-963),
// optionally fused with the PCG direction update  d <- M^-1 r + beta d  of MFEM's
// CGSolver::Mult (reference call site laghos_solver.cpp:388).
//
// One CTA = one batch of the host schedule (host/batch_plan.hpp): NB elements that form a brick of
// the element grid, all NC velocity components.  One launch per colour of the schedule.
//
//   phase 0  the quadrature data D of the batch's elements is requested first with one bulk-async
//            copy per element (cp.async.bulk.shared.global + mbarrier: 1728 contiguous bytes per
//            element at Q3Q2; no registers, no LSU wavefronts); meanwhile the batch's UNIQUE dofs
//            are loaded once, coalesced along lattice rows, into shared memory (with the fused
//            variant the new search direction is formed here and written back by the dof's
//            first-writer batch);
//   phase B  thread (element, component, qx): the whole yz sum factorisation of one qx plane in
//            registers -- x contraction straight from the unique-dof array (broadcast reads), y, z,
//            times D (from the shared slab), z^t, y^t -- 544 DFMA per thread at Q3Q2 with only
//            64 + 36 + 16 shared-memory accesses; 1D tables come from the kernel-parameter
//            constant bank, the thread's own B(qx,.) row lives in registers;
//   phase C  thread (element, component, dy, dz): x^t contraction of the six qx partials;
//   phase D  thread per unique dof: contributions of the batch's elements are summed in a fixed
//            order (CSR of the schedule) and written with a plain store (first writer of the dof)
//            or load-add-store (later colours).  griddepcontrol.wait in front of this phase is the
//            only dependency on the previous colour's launch: with programmatic dependent launch
//            everything before it overlaps the previous colour's tail.
//   The PCG denominator d^t A d = sum_q D u_q^2 is accumulated in phase B at no extra traffic.
#pragma once
#include "common.cuh"
#include "pcg.cuh"

namespace lagb {
namespace tuned {

struct BrickArgs
{
   int batch0, nbatch_launch;            // first batch of this launch (colour), number of batches
   int UP, NE;                           // padded unique capacity, element count
   int64_t cstride;                      // component stride of the L-vectors (byNODES)
   const int *belem; const int *bnuniq; const uint32_t *buid; const int *btab;
   const uint16_t *lidx, *uoff, *upos;   // index tables (lidx rows padded to NDP)
   const uint16_t *ucon;                 // [ntab][UP][8] plane slots of the contributions to a unique dof (0xffff = none)
   const double *Dq;                     // [NE*NQ]
   const double *x;                      // plain input (FUSE = false)
   const double *r, *dold; double *dnew; // fused PCG direction update (FUSE = true)
   const double *dinv; const unsigned char *ess; const pcg::State *st; int comp0;
   double *y;                            // output L-vector(s)
   double *den_part;                     // [nbatch*NC] partial d^t A d (WITH_DEN)
};

template<int D1D, int Q1D, int NB, int NC>
struct MassBrickCfg
{
   static constexpr int DD = D1D*D1D, QQ = Q1D*Q1D, ND = D1D*DD, NQ = Q1D*QQ;
   static constexpr int NDP = ((ND + 7)/8)*8;          // lidx row stride (16-byte vector loads)
   static constexpr int TB = NB*NC*Q1D;                // phase-B threads
   static constexpr int T = ((TB + 31)/32)*32;
   // per (element, component) slot of qx partials [dy + D1D*dz][qx]; stride = Q1D (mod 16) and even:
   // lanes (ec, qx) of a warp then hit consecutive 64-bit banks
   static constexpr int TS0 = Q1D*DD;
   static constexpr int TSP = TS0 + (((Q1D - TS0) % 16) + 16) % 16;
   static constexpr int NEC = NB*NC;
   static constexpr int ES = ND;                       // E staging stride per (element, component): [dx][dy + D1D*dz]
   static constexpr int R1 = (NB*NQ > NEC*ES) ? NB*NQ : NEC*ES;   // D slab, later E staging (doubles)
   static constexpr int KU = (NB*ND + T - 1)/T;        // unique-dof items per thread (upper bound)
   static constexpr int KC = (NEC*DD + T - 1)/T;       // phase-C items per thread
   static size_t smem_bytes(int UP) { return 16 + sizeof(double)*((size_t)R1 + (size_t)NC*UP + (size_t)NEC*TSP); }
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template<int D1D, int Q1D, int NB, int NC, bool WITH_DEN, bool FUSE, int MINB>
__global__ void __launch_bounds__((MassBrickCfg<D1D,Q1D,NB,NC>::T), MINB)
mass3d_brick(const __grid_constant__ DevTables<D1D,Q1D> tab, const __grid_constant__ BrickArgs a)
{
   using C = MassBrickCfg<D1D,Q1D,NB,NC>;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   uint64_t *mbar = reinterpret_cast<uint64_t*>(smem_raw);
   double *Ds = reinterpret_cast<double*>(smem_raw + 16);   // [NB][NQ], later E staging
   double *Xs = Ds + C::R1;                                  // [NC][UP]
   double *Ts = Xs + (size_t)NC*a.UP;                        // [NEC][TSP]
   const int t = threadIdx.x;
   const int kb = a.batch0 + blockIdx.x;                     // batch index in the schedule
   const int UP = a.UP;

   // let the next colour's launch start as early as possible (it blocks at its own phase D)
   asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
   if (FUSE) { if (a.st->all_done) { return; } }

   const int *el = a.belem + (size_t)kb*NB;
   int nel = 0;
#pragma unroll
   for (int j = 0; j < NB; j++) { nel += (__ldg(el + j) >= 0) ? 1 : 0; }

   // ---- phase 0a: bulk-async copies of the elements' quadrature data ----
   if (t == 0)
   {
      const uint32_t mb = smem_u32(mbar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"((uint32_t)(nel*C::NQ*sizeof(double))) : "memory");
      for (int j = 0; j < nel; j++)
      {
         const double *src = a.Dq + (size_t)__ldg(el + j)*C::NQ;
         asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                      :: "r"(smem_u32(Ds + (size_t)j*C::NQ)), "l"(src), "r"((uint32_t)(C::NQ*sizeof(double))), "r"(mb) : "memory");
      }
   }
   // ---- phase 0b: unique dofs -> shared memory ----
   const int nu = __ldg(a.bnuniq + kb);
   const uint32_t *uidp = a.buid + (size_t)kb*UP;
   uint32_t uw[C::KU];
   {
      double beta[NC]; bool skip[NC];
      if (FUSE)
      {
#pragma unroll
         for (int c = 0; c < NC; c++) { beta[c] = a.st->beta[c]; skip[c] = a.st->done[c] != 0; }
      }
#pragma unroll
      for (int k = 0; k < C::KU; k++)
      {
         const int u = t + k*C::T;
         uw[k] = (u < nu) ? __ldg(uidp + u) : 0u;
      }
#pragma unroll
      for (int k = 0; k < C::KU; k++)
      {
         const int u = t + k*C::T;
         if (u < nu)
         {
            const int64_t id = (int64_t)(uw[k] & 0x7fffffffu);
            if (FUSE)
            {
               const double di = a.dinv[id];
               const unsigned int em = (unsigned int)a.ess[id] >> a.comp0;
               const bool first = (uw[k] >> 31) != 0;
#pragma unroll
               for (int c = 0; c < NC; c++)
               {
                  const double rr = a.r[id + c*a.cstride], dv = a.dold[id + c*a.cstride];
                  const double zz = ((em >> c) & 1u) ? 0.0 : di*rr;
                  const double X = skip[c] ? dv : zz + beta[c]*dv;
                  Xs[c*UP + u] = X;
                  if (first) { a.dnew[id + c*a.cstride] = X; }
               }
            }
            else
            {
#pragma unroll
               for (int c = 0; c < NC; c++) { Xs[c*UP + u] = a.x[id + c*a.cstride]; }
            }
         }
      }
   }
   __syncthreads();   // Xs complete; mbarrier initialised for everyone
   // wait for the quadrature data (phase parity 0)
   {
      const uint32_t mb = smem_u32(mbar);
      uint32_t ok = 0;
      while (!ok)
      {
         asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                      : "=r"(ok) : "r"(mb) : "memory");
      }
   }
   const int tabid = __ldg(a.btab + kb);
   // ---- phase B: per (element, component, qx): x, y, z, D, z^t, y^t in registers ----
   double den = 0.0;
   {
      const int qx = t % Q1D, ec = t / Q1D;
      const int c = ec % NC, e = ec / NC;
      if (t < C::TB && e < nel)
      {
         double bq[D1D];
#pragma unroll
         for (int dx = 0; dx < D1D; dx++) { bq[dx] = tab.B[qx + Q1D*dx]; }
         const uint16_t *li = a.lidx + ((size_t)tabid*NB + e)*C::NDP;
         const double *Xc = Xs + c*UP;
         double V[Q1D][D1D];
#pragma unroll
         for (int dz = 0; dz < D1D; dz++)
         {
            double U[D1D];
            if constexpr (C::DD % 8 == 0)
            {
               // DD 16-bit slots of this dz slab: DD/8 vector loads (L1-resident table)
               uint32_t w[C::DD/2];
#pragma unroll
               for (int v = 0; v < C::DD/8; v++)
               {
                  const uint4 q = __ldg(reinterpret_cast<const uint4*>(li + dz*C::DD) + v);
                  w[4*v] = q.x; w[4*v + 1] = q.y; w[4*v + 2] = q.z; w[4*v + 3] = q.w;
               }
#pragma unroll
               for (int dy = 0; dy < D1D; dy++)
               {
                  double u = 0.0;
#pragma unroll
                  for (int dx = 0; dx < D1D; dx++)
                  {
                     const int i = dx + D1D*dy;
                     const uint32_t s = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu);
                     u += bq[dx]*Xc[s];
                  }
                  U[dy] = u;
               }
            }
            else
            {
#pragma unroll
               for (int dy = 0; dy < D1D; dy++)
               {
                  double u = 0.0;
#pragma unroll
                  for (int dx = 0; dx < D1D; dx++) { u += bq[dx]*Xc[__ldg(li + dx + D1D*(dy + D1D*dz))]; }
                  U[dy] = u;
               }
            }
#pragma unroll
            for (int qy = 0; qy < Q1D; qy++)
            {
               double v = 0.0;
#pragma unroll
               for (int dy = 0; dy < D1D; dy++) { v += tab.B[qy + Q1D*dy]*U[dy]; }
               V[qy][dz] = v;
            }
         }
         const double *De = Ds + (size_t)e*C::NQ + qx;
#pragma unroll
         for (int qy = 0; qy < Q1D; qy++)
         {
            double W[Q1D];
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++)
            {
               double w = 0.0;
#pragma unroll
               for (int dz = 0; dz < D1D; dz++) { w += tab.B[qz + Q1D*dz]*V[qy][dz]; }
               const double dw = De[Q1D*(qy + Q1D*qz)]*w;
               if (WITH_DEN) { den += dw*w; }
               W[qz] = dw;
            }
#pragma unroll
            for (int dz = 0; dz < D1D; dz++)
            {
               double v = 0.0;
#pragma unroll
               for (int qz = 0; qz < Q1D; qz++) { v += tab.B[qz + Q1D*dz]*W[qz]; }
               V[qy][dz] = v;
            }
         }
         double *tp = Ts + (size_t)ec*C::TSP + qx;
#pragma unroll
         for (int dz = 0; dz < D1D; dz++)
#pragma unroll
            for (int dy = 0; dy < D1D; dy++)
            {
               double u = 0.0;
#pragma unroll
               for (int qy = 0; qy < Q1D; qy++) { u += tab.B[qy + Q1D*dy]*V[qy][dz]; }
               tp[(dy + D1D*dz)*Q1D] = u;
            }
      }
   }
   __syncthreads();   // Ts complete, D slab dead
   // ---- phase C: x^t over the qx partials -> E staging (aliases the D slab) ----
   double *Es = Ds;
#pragma unroll
   for (int k = 0; k < C::KC; k++)
   {
      const int it = t + k*C::T;
      const int ec = it / C::DD, j = it - ec*C::DD;
      if (it < C::NEC*C::DD && ec / NC < nel)
      {
         const double *tp = Ts + (size_t)ec*C::TSP + j*Q1D;
         double tv[Q1D];
         if constexpr (Q1D % 2 == 0)
         {
#pragma unroll
            for (int q = 0; q < Q1D/2; q++)
            {
               const double2 v2 = reinterpret_cast<const double2*>(tp)[q];
               tv[2*q] = v2.x; tv[2*q + 1] = v2.y;
            }
         }
         else
         {
#pragma unroll
            for (int q = 0; q < Q1D; q++) { tv[q] = tp[q]; }
         }
#pragma unroll
         for (int dx = 0; dx < D1D; dx++)
         {
            double o = 0.0;
#pragma unroll
            for (int qx = 0; qx < Q1D; qx++) { o += tab.B[qx + Q1D*dx]*tv[qx]; }
            Es[(size_t)ec*C::ES + j + C::DD*dx] = o;
         }
      }
   }
   __syncthreads();
   // ---- phase D: fixed-order sum per unique dof, plain store / load-add-store ----
   // the only dependency on the previous colour's launch
   asm volatile("griddepcontrol.wait;" ::: "memory");
   {
      const uint16_t *uo = a.uoff + (size_t)tabid*(UP + 1);
      const uint16_t *up = a.upos + (size_t)tabid*NB*C::ND;
#pragma unroll
      for (int k = 0; k < C::KU; k++)
      {
         const int u = t + k*C::T;
         if (u < nu)
         {
            const int p0 = __ldg(uo + u), p1 = __ldg(uo + u + 1);
            double s[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) { s[c] = 0.0; }
            for (int p = p0; p < p1; p++)
            {
               const int pos = __ldg(up + p);
               const int e2 = pos / C::ND, i = pos - e2*C::ND;
               const int dx = i % D1D, j = i / D1D;
               const double *ep = Es + (size_t)e2*NC*C::ES + j + C::DD*dx;
#pragma unroll
               for (int c = 0; c < NC; c++) { s[c] += ep[c*C::ES]; }
            }
            const int64_t id = (int64_t)(uw[k] & 0x7fffffffu);
            if (uw[k] >> 31)
            {
#pragma unroll
               for (int c = 0; c < NC; c++) { a.y[id + c*a.cstride] = s[c]; }
            }
            else
            {
               double o[NC];
#pragma unroll
               for (int c = 0; c < NC; c++) { o[c] = __ldcg(a.y + id + c*a.cstride); }
#pragma unroll
               for (int c = 0; c < NC; c++) { a.y[id + c*a.cstride] = o[c] + s[c]; }
            }
         }
      }
   }
   if (WITH_DEN)
   {
      // deterministic block reduction of the thread partials (component c of thread t = (ec, qx))
      __syncthreads();
      double *red = Ts;
      constexpr int NW = C::T/32;
      const int cme = (t / Q1D) % NC;
#pragma unroll
      for (int c = 0; c < NC; c++)
      {
         double v = (cme == c) ? den : 0.0;
         for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
         if ((t & 31) == 0) { red[c*NW + (t >> 5)] = v; }
      }
      __syncthreads();
      if (t < NC)
      {
         double s = 0.0;
         for (int w = 0; w < NW; w++) { s += red[t*NW + w]; }
         a.den_part[(size_t)kb*NC + t] = s;
      }
   }
}

// ---------------------------------------------------------------------------------------------
// Second brick kernel: the slice / column / slice register contractions of device/mass3d.cuh
// (every thread busy in every phase, 4*D1D*Q1D^2 doubles of shared-memory traffic per element
// and component) on the deduplicated, coalesced brick gather / fixed-order store of the schedule.
//
//   phase 0  D of the batch's elements: one bulk-async copy per element into shared memory
//            (DBULK: cp.async.bulk + mbarrier) or prefetched into registers; unique dofs -> Xs
//            (FUSE: the PCG direction update is formed here, first writer stores it);
//   phase A  thread (c, e, dz): slice values from Xs through the 16-bit slot table, x then y
//            contraction in registers, Q1D^2 plane -> shared memory;
//   phase B  thread per quadrature column (all NC components, D read once): z, *D, z^t in place;
//   phase C  thread (c, e, dz): y^t, x^t; the slice's D1D^2 results overwrite the head of its plane;
//   phase D  thread per unique dof: fixed-order sum of its <= 8 contributions (fixed-width table),
//            plain store (first writer) or load-add-store; griddepcontrol.wait in front.
// ---------------------------------------------------------------------------------------------
template<int D1D, int Q1D, int NB, int NC>
struct MassBrick2Cfg
{
   static constexpr int DD = D1D*D1D, QQ = Q1D*Q1D, ND = D1D*DD, NQ = Q1D*QQ;
   static constexpr int NDP = ((ND + 7)/8)*8;
   static constexpr int TG = NB*D1D, TA = NC*TG, T = ((TA + 31)/32)*32;
   static constexpr int NCOL = (NB*QQ + T - 1)/T;
   static constexpr int PLANE = ((QQ > DD ? QQ : DD) | 1);
   static constexpr int CPL = NB*D1D*PLANE;            // plane doubles per component
   static constexpr int KU = (NB*ND + T - 1)/T;
   static constexpr int MAXCON = 8;                    // contributions per unique dof inside one batch
   static size_t smem_bytes(int UP, bool dbulk)
   { return 16 + sizeof(double)*((dbulk ? (size_t)NB*NQ : 0) + (size_t)NC*UP + (size_t)NC*CPL); }
};

template<int D1D, int Q1D, int NB, int NC, bool WITH_DEN, bool FUSE, int MINB, bool DBULK>
__global__ void __launch_bounds__((MassBrick2Cfg<D1D,Q1D,NB,NC>::T), MINB)
mass3d_brick2(const __grid_constant__ DevTables<D1D,Q1D> tab, const __grid_constant__ BrickArgs a)
{
   using C = MassBrick2Cfg<D1D,Q1D,NB,NC>;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   uint64_t *mbar = reinterpret_cast<uint64_t*>(smem_raw);
   double *Ds = reinterpret_cast<double*>(smem_raw + 16);         // [NB][NQ] (DBULK)
   double *Xs = Ds + (DBULK ? NB*C::NQ : 0);                      // [NC][UP]
   const int UP = a.UP;
   double *sV = Xs + (size_t)NC*UP;                               // [c][e][dz][PLANE]
   const int t = threadIdx.x;
   const int kb = a.batch0 + blockIdx.x;

   asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
   if (FUSE) { if (a.st->all_done) { return; } }

   const int *el = a.belem + (size_t)kb*NB;
   int nel = 0;
#pragma unroll
   for (int j = 0; j < NB; j++) { nel += (__ldg(el + j) >= 0) ? 1 : 0; }
   const int ncols = nel*C::QQ;

   // ---- phase 0a: quadrature data ----
   double dq[DBULK ? 1 : C::NCOL][Q1D];
   if (DBULK)
   {
      if (t == 0)
      {
         const uint32_t mb = smem_u32(mbar);
         asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb));
         asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
         asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"((uint32_t)(nel*C::NQ*sizeof(double))) : "memory");
         for (int j = 0; j < nel; j++)
         {
            const double *src = a.Dq + (size_t)__ldg(el + j)*C::NQ;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(Ds + (size_t)j*C::NQ)), "l"(src), "r"((uint32_t)(C::NQ*sizeof(double))), "r"(mb) : "memory");
         }
      }
   }
   else
   {
#pragma unroll
      for (int k = 0; k < C::NCOL; k++)
      {
         const int f = t + k*C::T;
         if (f < ncols)
         {
            const int e2 = f / C::QQ, col = f - e2*C::QQ;
            const double *dptr = a.Dq + (size_t)__ldg(el + e2)*C::NQ + col;
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++) { dq[k][qz] = __ldg(dptr + C::QQ*qz); }
         }
      }
   }
   // ---- phase 0b: unique dofs -> shared memory ----
   const int nu = __ldg(a.bnuniq + kb);
   const uint32_t *uidp = a.buid + (size_t)kb*UP;
   uint32_t uw[C::KU];
   {
      double beta[NC]; bool skip[NC];
      if (FUSE)
      {
#pragma unroll
         for (int c = 0; c < NC; c++) { beta[c] = a.st->beta[c]; skip[c] = a.st->done[c] != 0; }
      }
#pragma unroll
      for (int k = 0; k < C::KU; k++)
      {
         const int u = t + k*C::T;
         uw[k] = (u < nu) ? __ldg(uidp + u) : 0u;
      }
      // later colours read-modify-write y: pull those lines into L2 now (L2 is the coherence point,
      // so this is safe before griddepcontrol.wait)
#pragma unroll
      for (int k = 0; k < C::KU; k++)
      {
         const int u = t + k*C::T;
         if (u < nu && (uw[k] >> 31) == 0)
         {
            const int64_t id = (int64_t)(uw[k] & 0x7fffffffu);
#pragma unroll
            for (int cc = 0; cc < NC; cc++) { asm volatile("prefetch.global.L2 [%0];" :: "l"(a.y + id + cc*a.cstride)); }
         }
      }
      if (FUSE)
      {
#pragma unroll
         for (int k = 0; k < C::KU; k++)
         {
            const int u = t + k*C::T;
            if (u < nu)
            {
               const int64_t id = (int64_t)(uw[k] & 0x7fffffffu);
               const double di = a.dinv[id];
               const unsigned int em = (unsigned int)a.ess[id] >> a.comp0;
               const bool first = (uw[k] >> 31) != 0;
               double rr[NC], dv[NC];
#pragma unroll
               for (int c = 0; c < NC; c++) { rr[c] = a.r[id + c*a.cstride]; dv[c] = a.dold[id + c*a.cstride]; }
#pragma unroll
               for (int c = 0; c < NC; c++)
               {
                  const double zz = ((em >> c) & 1u) ? 0.0 : di*rr[c];
                  const double X = skip[c] ? dv[c] : zz + beta[c]*dv[c];
                  Xs[c*UP + u] = X;
                  if (first) { a.dnew[id + c*a.cstride] = X; }
               }
            }
         }
      }
      else
      {
         double xv[C::KU][NC];
#pragma unroll
         for (int k = 0; k < C::KU; k++)
         {
            const int u = t + k*C::T;
            const int64_t id = (int64_t)(uw[k] & 0x7fffffffu);
#pragma unroll
            for (int c = 0; c < NC; c++) { xv[k][c] = (u < nu) ? a.x[id + c*a.cstride] : 0.0; }
         }
#pragma unroll
         for (int k = 0; k < C::KU; k++)
         {
            const int u = t + k*C::T;
            if (u < nu)
            {
#pragma unroll
               for (int c = 0; c < NC; c++) { Xs[c*UP + u] = xv[k][c]; }
            }
         }
      }
   }
   __syncthreads();   // Xs complete; mbarrier initialised for everyone
   const int tabid = __ldg(a.btab + kb);
   const int c = t / C::TG, r = t - c*C::TG;
   const int e_loc = r / D1D, dz = r % D1D;
   const bool active = (t < C::TA) && (e_loc < nel);
   double *pl = sV + (size_t)c*C::CPL + (size_t)(e_loc*D1D + dz)*C::PLANE;
   // ---- phase A: slice values from Xs, x then y contraction, plane -> shared memory ----
   if (active)
   {
      const uint16_t *li = a.lidx + ((size_t)tabid*NB + e_loc)*C::NDP + dz*C::DD;
      const double *Xc = Xs + c*UP;
      double XG[C::DD];
      if constexpr (C::DD % 8 == 0)
      {
#pragma unroll
         for (int v = 0; v < C::DD/8; v++)
         {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(li) + v);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) { XG[8*v + 2*j] = Xc[w[j] & 0xffffu]; XG[8*v + 2*j + 1] = Xc[w[j] >> 16]; }
         }
      }
      else
      {
#pragma unroll
         for (int i = 0; i < C::DD; i++) { XG[i] = Xc[__ldg(li + i)]; }
      }
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
   if (DBULK)
   {
      const uint32_t mb = smem_u32(mbar);
      uint32_t ok = 0;
      while (!ok)
      {
         asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                      : "=r"(ok) : "r"(mb) : "memory");
      }
   }
   // ---- phase B: z contraction, scale by D, z back (flat element-column index, all components) ----
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
         double dl[Q1D];
#pragma unroll
         for (int qz = 0; qz < Q1D; qz++) { dl[qz] = DBULK ? Ds[(size_t)e2*C::NQ + col + C::QQ*qz] : dq[DBULK ? 0 : k][qz]; }
#pragma unroll
         for (int cc = 0; cc < NC; cc++)
         {
            double *colp = sV + (size_t)cc*C::CPL + (size_t)(e2*D1D)*C::PLANE + col;
            double V[D1D], W[Q1D];
#pragma unroll
            for (int kk = 0; kk < D1D; kk++) { V[kk] = colp[kk*C::PLANE]; }
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++)
            {
               double w = 0.0;
#pragma unroll
               for (int kk = 0; kk < D1D; kk++) { w += tab.B[qz + Q1D*kk]*V[kk]; }
               const double dw = dl[qz]*w;
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
   __syncthreads();
   // ---- phase D: fixed-order sum per unique dof; the only dependency on the previous colour ----
   {
      const uint4 *ucp = reinterpret_cast<const uint4*>(a.ucon) + (size_t)tabid*UP;
      uint4 qc[C::KU];
#pragma unroll
      for (int k = 0; k < C::KU; k++)
      {
         const int u = t + k*C::T;
         qc[k] = (u < nu) ? __ldg(ucp + u) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");
      // all read-modify-write loads of the thread first (independent requests in flight)
      double o[C::KU][NC];
#pragma unroll
      for (int k = 0; k < C::KU; k++)
      {
         const int u = t + k*C::T;
         const int64_t id = (int64_t)(uw[k] & 0x7fffffffu);
         const bool rmw = (u < nu) && (uw[k] >> 31) == 0;
#pragma unroll
         for (int cc = 0; cc < NC; cc++) { o[k][cc] = rmw ? __ldcg(a.y + id + cc*a.cstride) : 0.0; }
      }
#pragma unroll
      for (int k = 0; k < C::KU; k++)
      {
         const int u = t + k*C::T;
         if (u < nu)
         {
            const int64_t id = (int64_t)(uw[k] & 0x7fffffffu);
            const uint32_t w[4] = {qc[k].x, qc[k].y, qc[k].z, qc[k].w};
            double s[NC];
#pragma unroll
            for (int cc = 0; cc < NC; cc++) { s[cc] = 0.0; }
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
               const uint32_t p0 = w[j] & 0xffffu, p1 = w[j] >> 16;
               if (p0 != 0xffffu)
               {
#pragma unroll
                  for (int cc = 0; cc < NC; cc++) { s[cc] += sV[(size_t)cc*C::CPL + p0]; }
               }
               if (p1 != 0xffffu)
               {
#pragma unroll
                  for (int cc = 0; cc < NC; cc++) { s[cc] += sV[(size_t)cc*C::CPL + p1]; }
               }
            }
            // first writer: o = 0 and 0 + s = s exactly
#pragma unroll
            for (int cc = 0; cc < NC; cc++) { a.y[id + cc*a.cstride] = o[k][cc] + s[cc]; }
         }
      }
   }
   if (WITH_DEN)
   {
      __syncthreads();
      double *red = sV;
      constexpr int NW = C::T/32;
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
         a.den_part[(size_t)kb*NC + t] = s;
      }
   }
}

} // namespace tuned
} // namespace lagb
