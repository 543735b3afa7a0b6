// Persistent, dataflow-ordered brick H1 mass apply:  y = G^t B^t D B G x  in ONE launch, no atomics
// on the data path, bitwise reproducible
// (reference MassPAOperator::Mult -> MFEM MassIntegrator::AddMultPA, laghos_assembly.cpp:117-121;
//  arithmetic as in amr/laghos_assembly.cpp:878-963).
//
// The coloured schedule of host/batch_plan.hpp fixes the summation order of every shared dof
// (colour order).  The first two brick kernels run one launch per colour; here the colour
// boundaries are replaced by per-batch completion flags:
//
//   * batches are handed out in schedule order by a global work counter (dynamic: only RUNNING
//     CTAs hold batches, so a batch's lower-coloured neighbours are always finished or running --
//     no deadlock, whatever the residency; automatic load balance, one tail instead of eight);
//   * before its read-modify-write phase a batch waits for the flags of the lower-coloured batches
//     that share a dof with it (<= 26 on a brick grid; almost always already set), and publishes
//     its own flag after its stores (bar.sync + one gpu-scope fence + flag store);
//   * the CTA is persistent, so the loads of the NEXT batch are in flight while the current one
//     computes: unique dofs by cp.async (LDGSTS) into the dead Xs buffer after phase A, the
//     quadrature data D into registers (or by cp.async.bulk + mbarrier into the dead D slab) after
//     phase B, schedule metadata one batch ahead.  No global-memory latency on the critical path.
//
// Phases A-D are those of mass3d_brick2 (device/mass3d_brick.cuh).
#pragma once
#include "mass3d_brick.cuh"

namespace lagb {
namespace tuned {

struct Brick3Args
{
   BrickArgs b;
   int nbatch_total;
   const int4 *bmeta;        // [nbatch] {nel, nuniq, table id, unused}
   const int *deps;          // [nbatch][32] lower-coloured batches sharing a dof, -1 = none
   int *flags;               // [nbatch] == epoch once the batch's stores are visible
   int *work_ctr;            // next batch to hand out (zeroed before the launch)
   int epoch;
};

template<int D1D, int Q1D, int NB, int NC>
struct MassBrick3Cfg : MassBrick2Cfg<D1D,Q1D,NB,NC>
{
   using Base = MassBrick2Cfg<D1D,Q1D,NB,NC>;
   static constexpr int CTL = (32 + NC*8*(Base::T/32) + 127)/128*128;   // control block: mbarrier, next-batch slot, den scratch
   static size_t smem_bytes(int UP, bool dbulk)
   { return CTL + sizeof(double)*((dbulk ? (size_t)NB*Base::NQ : 0) + (size_t)NC*UP + (size_t)NC*Base::CPL); }
};

// schedule metadata loads that must stay where they are written (one batch ahead of their use)
__device__ __forceinline__ int ld_nc_i32(const int *p)
{
   int v; asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ int4 ld_nc_i32x4(const int4 *p)
{
   int4 v; asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v;
}
__device__ __forceinline__ int ld_volatile_i32(const int *p)
{
   int v; asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

template<int D1D, int Q1D, int NB, int NC, bool WITH_DEN, int MINB, bool DBULK>
__global__ void __launch_bounds__((MassBrick2Cfg<D1D,Q1D,NB,NC>::T), MINB)
mass3d_brick3(const __grid_constant__ DevTables<D1D,Q1D> tab, const __grid_constant__ Brick3Args g)
{
   using C = MassBrick3Cfg<D1D,Q1D,NB,NC>;
   const BrickArgs &a = g.b;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   uint64_t *mbar = reinterpret_cast<uint64_t*>(smem_raw);
   int *s_next = reinterpret_cast<int*>(smem_raw + 16);
   double *s_red = reinterpret_cast<double*>(smem_raw + 32);       // [NC][NW]
   double *Ds = reinterpret_cast<double*>(smem_raw + C::CTL);      // [NB][NQ] (DBULK)
   double *Xs = Ds + (DBULK ? NB*C::NQ : 0);                       // [NC][UP]
   const int UP = a.UP;
   double *sV = Xs + (size_t)NC*UP;                                // [c][e][dz][PLANE]
   const int t = threadIdx.x;
   constexpr int NW = C::T/32;
   static_assert(NC*NW*8 + 32 <= C::CTL, "den scratch does not fit the control block");

   const int c = t / C::TG, r = t - c*C::TG;
   const int e_loc = r / D1D, dz = r % D1D;
   double *pl = sV + (size_t)c*C::CPL + (size_t)(e_loc*D1D + dz)*C::PLANE;

   if (t == 0)
   {
      s_next[0] = atomicAdd(g.work_ctr, 1);
      if (DBULK)
      {
         asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar)));
         asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
   }
   __syncthreads();
   int k = s_next[0];
   if (k >= g.nbatch_total) { return; }
   __syncthreads();
   if (t == 0) { s_next[0] = atomicAdd(g.work_ctr, 1); }   // the batch after k

   // ---- helpers (all inlined) ----
   auto load_uids = [&](int kb, int nu, uint32_t *uw)
   {
      const uint32_t *uidp = a.buid + (size_t)kb*UP;
#pragma unroll
      for (int q = 0; q < C::KU; q++) { const int u = t + q*C::T; uw[q] = (u < nu) ? __ldg(uidp + u) : 0u; }
   };
   auto issue_x = [&](int nu, const uint32_t *uw)
   {
#pragma unroll
      for (int q = 0; q < C::KU; q++)
      {
         const int u = t + q*C::T;
         if (u < nu)
         {
            const int64_t id = (int64_t)(uw[q] & 0x7fffffffu);
#pragma unroll
            for (int cc = 0; cc < NC; cc++)
            {
               asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem_u32(Xs + cc*UP + u)), "l"(a.x + id + cc*a.cstride) : "memory");
            }
            if ((uw[q] >> 31) == 0)
            {
               // read-modify-write target of a later colour: pull the line into L2 now
#pragma unroll
               for (int cc = 0; cc < NC; cc++) { asm volatile("prefetch.global.L2 [%0];" :: "l"(a.y + id + cc*a.cstride)); }
            }
         }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
   };
   double dq[DBULK ? 1 : C::NCOL][Q1D];
   // element ids of the thread's phase-B columns (registers, loaded one batch ahead)
   auto load_elc = [&](int kb, int nel, int *elc)
   {
      const int *el = a.belem + (size_t)kb*NB;
#pragma unroll
      for (int q = 0; q < C::NCOL; q++)
      {
         const int f = t + q*C::T;
         elc[q] = (f < nel*C::QQ) ? ld_nc_i32(el + f / C::QQ) : -1;
      }
   };
   auto issue_D = [&](int kb, int nel, const int *elc)
   {
      if (DBULK)
      {
         if (t == 0)
         {
            const int *el = a.belem + (size_t)kb*NB;
            const uint32_t mb = smem_u32(mbar);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the slab are done (bar.sync before)
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
         for (int q = 0; q < C::NCOL; q++)
         {
            const int f = t + q*C::T;
            if (elc[q] >= 0)
            {
               const int e2 = f / C::QQ, col = f - e2*C::QQ;
               const double *dptr = a.Dq + (size_t)elc[q]*C::NQ + col;
#pragma unroll
               for (int qz = 0; qz < Q1D; qz++) { dq[DBULK ? 0 : q][qz] = __ldg(dptr + C::QQ*qz); }
            }
         }
      }
   };

   // ---- prologue: first batch ----
   int4 meta = ld_nc_i32x4(g.bmeta + k);
   int dep = (t < 32) ? ld_nc_i32(g.deps + (size_t)k*32 + t) : -1;
   uint32_t uw[C::KU];
   load_uids(k, meta.y, uw);
   issue_x(meta.y, uw);
   {
      int elc[C::NCOL];
      load_elc(k, meta.x, elc);
      issue_D(k, meta.x, elc);
   }
   uint32_t dpar = 0;
   asm volatile("cp.async.wait_all;" ::: "memory");
   __syncthreads();                            // Xs of the first batch complete; s_next holds the second batch

   while (true)
   {
      const int nel = meta.x, nu = meta.y, tabid = meta.z;
      const int ncols = nel*C::QQ;
      const bool active = (t < C::TA) && (e_loc < nel);
      const int kn = s_next[0];                // written after S2 of the previous batch (or in the prologue)
      const bool have_n = kn < g.nbatch_total;
      // schedule metadata of the next batch (registers; consumed after phase A)
      int4 meta_n = make_int4(0, 0, 0, 0);
      int dep_n = -1;
      uint32_t uw_n[C::KU];
      int elc_n[C::NCOL];
      int nxt2 = g.nbatch_total;
      if (have_n)
      {
         meta_n = ld_nc_i32x4(g.bmeta + kn);
         dep_n = (t < 32) ? ld_nc_i32(g.deps + (size_t)kn*32 + t) : -1;
         if (t == 0) { nxt2 = atomicAdd(g.work_ctr, 1); }   // the batch after kn (published after S2)
      }
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
      if (have_n) { load_uids(kn, meta_n.y, uw_n); load_elc(kn, meta_n.x, elc_n); }
      __syncthreads();                         // S2: planes complete; Xs dead; everyone has read s_next
      if (t == 0) { s_next[0] = nxt2; }
      if (have_n) { issue_x(meta_n.y, uw_n); }
      if (DBULK)
      {
         const uint32_t mb = smem_u32(mbar);
         uint32_t ok = 0;
         while (!ok)
         {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(mb), "r"(dpar) : "memory");
         }
         dpar ^= 1u;
      }
      // ---- phase B: z contraction, scale by D, z back (flat element-column index, all components) ----
      double den[NC];
#pragma unroll
      for (int cc = 0; cc < NC; cc++) { den[cc] = 0.0; }
#pragma unroll
      for (int q = 0; q < C::NCOL; q++)
      {
         const int f = t + q*C::T;
         if (f < ncols)
         {
            const int e2 = f / C::QQ, col = f - e2*C::QQ;
            double dl[Q1D];
#pragma unroll
            for (int qz = 0; qz < Q1D; qz++) { dl[qz] = DBULK ? Ds[(size_t)e2*C::NQ + col + C::QQ*qz] : dq[DBULK ? 0 : q][qz]; }
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
      if (WITH_DEN)
      {
#pragma unroll
         for (int cc = 0; cc < NC; cc++)
         {
            double v = den[cc];
            for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
            if ((t & 31) == 0) { s_red[cc*NW + (t >> 5)] = v; }
         }
      }
      __syncthreads();                         // S3: planes hold the z^t results; D of batch k dead
      if (have_n) { issue_D(kn, meta_n.x, elc_n); }
      if (WITH_DEN && t < NC)
      {
         double s = 0.0;
         for (int w = 0; w < NW; w++) { s += s_red[t*NW + w]; }
         a.den_part[(size_t)k*NC + t] = s;
      }
      // flags of the lower-coloured neighbours: read now, checked after phase C
      int fl = g.epoch;
      if (t < 32 && dep >= 0) { fl = ld_volatile_i32(g.flags + dep); }
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
      if (t < 32)
      {
         // dataflow wait (warp 0): every lower-coloured sharer has published its stores
         while (!__all_sync(0xffffffffu, fl == g.epoch))
         {
            if (dep >= 0 && fl != g.epoch) { fl = ld_volatile_i32(g.flags + dep); }
         }
         // no acquire fence: the read-modify-write loads below are issued after the bar.sync that follows this
         // loop and go to L2 (ld.global.cg), the point of coherence; nothing of y is ever cached in L1
      }
      __syncthreads();                         // S4: slice results complete, dependencies satisfied
      // ---- phase D: fixed-order sum per unique dof ----
      {
         const uint4 *ucp = reinterpret_cast<const uint4*>(a.ucon) + (size_t)tabid*UP;
         uint4 qc[C::KU];
         double o[C::KU][NC];
#pragma unroll
         for (int q = 0; q < C::KU; q++)
         {
            const int u = t + q*C::T;
            qc[q] = (u < nu) ? __ldg(ucp + u) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            const int64_t id = (int64_t)(uw[q] & 0x7fffffffu);
            const bool rmw = (u < nu) && (uw[q] >> 31) == 0;
#pragma unroll
            for (int cc = 0; cc < NC; cc++) { o[q][cc] = rmw ? __ldcg(a.y + id + cc*a.cstride) : 0.0; }
         }
#pragma unroll
         for (int q = 0; q < C::KU; q++)
         {
            const int u = t + q*C::T;
            if (u < nu)
            {
               const int64_t id = (int64_t)(uw[q] & 0x7fffffffu);
               const uint32_t w[4] = {qc[q].x, qc[q].y, qc[q].z, qc[q].w};
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
#pragma unroll
               for (int cc = 0; cc < NC; cc++) { a.y[id + cc*a.cstride] = o[q][cc] + s[cc]; }
            }
         }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();                         // S5: stores issued, plane heads read, Xs(kn) landed
      if (t == 0)
      {
         // release: the CTA's stores (ordered before this thread by bar.sync, cumulativity) become visible at
         // gpu scope before the flag.  st.release.gpu = MEMBAR.ALL.GPU + STG.STRONG: unlike __threadfence() it
         // does not invalidate L1 (CCTL.IVALL), so the index tables of the co-resident CTAs stay cached
         asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(g.flags + k), "r"(g.epoch) : "memory");
      }
      if (!have_n) { break; }
      k = kn; meta = meta_n; dep = dep_n;
#pragma unroll
      for (int q = 0; q < C::KU; q++) { uw[q] = uw_n[q]; }
   }
}

} // namespace tuned
} // namespace lagb
