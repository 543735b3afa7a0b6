// Launchers for the tuned 3D kernels (sm_100a).
#include "ctx.hpp"
#include <algorithm>
#include "device/mass3d.cuh"
#include "device/mass3d_brick.cuh"
#include "device/mass3d_brick3.cuh"
#include "device/staged3d.cuh"
#include "device/mass3d_pencil.cuh"

namespace lagb {

// opt-in dynamic shared memory: the attribute is per device, so it is tracked per context (one
// context = one device), not per process
template<typename K>
static int set_smem(Ctx &c, K kern, size_t bytes)
{
   size_t &have = c.smem_optin[(const void*)kern];
   if (have >= bytes && have != 0) { return LAGB_OK; }
   LAGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
   have = bytes;
   return LAGB_OK;
}

// NTQ and NTF must be multiples of 32 (full-warp shuffles).
// <D1D, Q1D, NB1, MINB1, NB3, MINB3, NTQ, NBF, NTF>: elements per CTA and minimum resident CTAs
// for the 1- and 3-component mass apply, threads per element of QUpdate, elements per CTA and
// threads of Force / Force^T.
template<int D1D, int Q1D, int NB1, int MINB1, int NB3, int MINB3, int NTQ, int NBF, int NTF>
struct TunedLaunch3D
{
   using Tab = DevTables<D1D,Q1D>;
   // elements per batch of the brick mass apply (orders without tuning variants)
   static constexpr int NBB = (D1D <= 2) ? 32 : (D1D == 3) ? 16 : (D1D <= 5) ? 8 : 4;
   static const Tab &tab(Ctx &c) { return *reinterpret_cast<const Tab*>(c.tab_blob.data()); }

   template<int NC, bool WITH_DEN, int NB, int MINB, bool DS = false, bool DG = false>
   static int mass_launch_v(Ctx &c, const double *x, double *y)
   {
      using Cfg = tuned::Mass3DCfg<D1D,Q1D,NB,NC>;
      auto kern = tuned::mass3d<D1D,Q1D,NB,NC,WITH_DEN,MINB,DS,DG>;
      { int rc = set_smem(c, kern, Cfg::SMEM_BYTES); if (rc) { return rc; } }
      const int grid = (c.NE + NB - 1)/NB;
      if (WITH_DEN && grid*NC > c.part_cap) { set_error("mass3d: partial buffer too small"); return LAGB_ERR_STATE; }
      LAGB_LAUNCH_K(c, kern, grid, Cfg::T, Cfg::SMEM_BYTES, tab(c), c.NE, (int64_t)c.ndofs, (const int*)c.d_map, (const double*)c.d_massD, x, y, c.d_part);
      if (WITH_DEN) { c.dt_nblocks = grid; }
      return LAGB_OK;
   }
   template<int NC, bool WITH_DEN>
   static int mass_launch(Ctx &c, const double *x, double *y)
   {
      if constexpr (NC == 3 && D1D == 4)   // tuning variants of the dominant kernel (lagb_tune_set key 0)
      {
         switch (c.tune[0])
         {
            case 1: return mass_launch_v<NC,WITH_DEN,8,6,true>(c, x, y);
            case 2: return mass_launch_v<NC,WITH_DEN,16,3,true,true>(c, x, y);
            case 3: return mass_launch_v<NC,WITH_DEN,8,5,true>(c, x, y);
            case 4: return mass_launch_v<NC,WITH_DEN,16,3,true>(c, x, y);
         }
         return mass_launch_v<NC,WITH_DEN,8,6,true,true>(c, x, y);   // measured best on B200 (profiles/microbench_r1_variants*.txt): 351 us
      }
      if constexpr (NC == 1 && D1D == 4)   // single-component apply (lagb_tune_set key 3)
      {
         switch (c.tune[3])
         {
            case 1: return mass_launch_v<NC,WITH_DEN,32,2,true>(c, x, y);
            case 2: return mass_launch_v<NC,WITH_DEN,16,4,true>(c, x, y);
            case 3: return mass_launch_v<NC,WITH_DEN,16,8,true>(c, x, y);
            case 4: return mass_launch_v<NC,WITH_DEN,32,2>(c, x, y);
         }
         return mass_launch_v<NC,WITH_DEN,8,8,true,true>(c, x, y);
      }
      if constexpr (D1D >= 5)
      {
         // pencil kernel (device/mass3d_pencil.cuh) against the slice kernel, measured on box01_hex -rs 4 (us):
         //   Q4Q3  1 comp 165 / 284   3 comp 345 / 343        Q5Q4  1 comp 880 / 417   3 comp 1574 / 1142
         // -> pencils only for the single-component apply at Q4Q3; lagb_tune_set key 0 = 5..7 selects them elsewhere
         constexpr int NBP = (NC == 1) ? ((D1D == 5) ? 8 : 4) : ((D1D == 5) ? 4 : 2);
         switch (c.tune[0])
         {
            case 5: return pencil_launch<NC,WITH_DEN,NBP,256>(c, x, y);
            case 6: return pencil_launch<NC,WITH_DEN,(NBP > 1 ? NBP/2 : 1),256>(c, x, y);
            case 7: return pencil_launch<NC,WITH_DEN,NBP*2,256>(c, x, y);
         }
         if constexpr (D1D == 5 && NC == 1) { return pencil_launch<NC,WITH_DEN,NBP,256>(c, x, y); }
      }
      if constexpr (D1D == 3 && NC == 3)   // Q2Q1 tuning variants (lagb_tune_set key 0)
      {
         switch (c.tune[0])
         {
            case 1: return mass_launch_v<NC,WITH_DEN,16,2,true,true>(c, x, y);
            case 2: return mass_launch_v<NC,WITH_DEN,64,1,true,true>(c, x, y);
            case 3: return mass_launch_v<NC,WITH_DEN,32,2,true,true>(c, x, y);
            case 4: return mass_launch_v<NC,WITH_DEN,16,4,true,true>(c, x, y);
            case 5: return mass_launch_v<NC,WITH_DEN,8,8,true,true>(c, x, y);
         }
      }
      // direct gather / scatter for the low orders; staged through shared memory where registers are tight
      return mass_launch_v<NC,WITH_DEN,(NC == 1) ? NB1 : NB3,(NC == 1) ? MINB1 : MINB3,(D1D <= 3),(D1D <= 3)>(c, x, y);
   }
   template<int NC, bool WITH_DEN, int NB, int NT>
   static int pencil_launch(Ctx &c, const double *x, double *y)
   {
      using Cfg = tuned::MassPencilCfg<D1D,Q1D,NC>;
      auto kern = tuned::mass3d_pencil<D1D,Q1D,NB,NC,NT,WITH_DEN>;
      constexpr size_t bytes = sizeof(double)*(size_t)NB*NC*Cfg::PER_EC;
      if (bytes > 227*1024) { set_error("mass3d_pencil: this launch variant does not fit shared memory at this order"); return LAGB_ERR_INVALID; }
      { int rc = set_smem(c, kern, bytes); if (rc) { return rc; } }
      const int grid = (c.NE + NB - 1)/NB;
      if (WITH_DEN && grid*NC > c.part_cap) { set_error("mass3d_pencil: partial buffer too small"); return LAGB_ERR_STATE; }
      LAGB_LAUNCH_K(c, kern, grid, NT, bytes, tab(c), c.NE, (int64_t)c.ndofs, (const int*)c.d_map, (const double*)c.d_massD, x, y, c.d_part);
      if (WITH_DEN) { c.dt_nblocks = grid; }
      return LAGB_OK;
   }
   // ---- brick schedule (device/mass3d_brick.cuh): one launch per colour, programmatic dependent launch ----
   template<int NC, bool WITH_DEN, bool FUSE, int NB, int MINB>
   static int brick_launch_v(Ctx &c, const MassBrickIn &in, double *y)
   {
      using Cfg = tuned::MassBrickCfg<D1D,Q1D,NB,NC>;
      DevPlan *pl = nullptr;
      int rc = get_plan(c, NB, &pl); if (rc) { return rc; }
      auto kern = tuned::mass3d_brick<D1D,Q1D,NB,NC,WITH_DEN,FUSE,MINB>;
      const size_t bytes = Cfg::smem_bytes(pl->UP);
      rc = set_smem(c, kern, bytes); if (rc) { return rc; }
      if (WITH_DEN && pl->nbatch*NC > c.part_cap) { set_error("mass3d_brick: partial buffer too small"); return LAGB_ERR_STATE; }
      tuned::BrickArgs a;
      a.UP = pl->UP; a.NE = c.NE; a.cstride = c.ndofs;
      a.belem = pl->belem; a.bnuniq = pl->bnuniq; a.buid = pl->buid; a.btab = pl->btab;
      a.lidx = pl->lidx; a.uoff = pl->uoff; a.upos = pl->upos;
      a.Dq = c.d_massD; a.x = in.x; a.r = in.r; a.dold = in.dold; a.dnew = in.dnew;
      a.dinv = c.d_dinv; a.ess = c.d_essmask; a.st = c.d_state; a.comp0 = in.comp0;
      a.y = y; a.den_part = c.d_part;
      for (int col = 0; col < pl->ncolors; col++)
      {
         a.batch0 = pl->color_begin[col]; a.nbatch_launch = pl->color_begin[col + 1] - a.batch0;
         if (a.nbatch_launch <= 0) { continue; }
         cudaLaunchConfig_t cfg = {};
         cfg.gridDim = dim3((unsigned)a.nbatch_launch); cfg.blockDim = dim3(Cfg::T); cfg.dynamicSmemBytes = bytes; cfg.stream = c.stream;
         cudaLaunchAttribute at[1];
         at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
         at[0].val.programmaticStreamSerializationAllowed = 1;
         cfg.attrs = at; cfg.numAttrs = (col > 0 && c.tune[5] == 0) ? 1 : 0;   // colour 0 waits for everything before it
         LAGB_CUDA(cudaLaunchKernelEx(&cfg, kern, tab(c), a));
         LAGB_LAUNCH_CHECK();
      }
      if (WITH_DEN) { c.dt_nblocks = pl->nbatch; }
      return LAGB_OK;
   }
   // second brick kernel (slice/column/slice compute on the deduplicated gather)
   template<int NC, bool WITH_DEN, bool FUSE, int NB, int MINB, bool DBULK>
   static int brick2_launch_v(Ctx &c, const MassBrickIn &in, double *y)
   {
      using Cfg = tuned::MassBrick2Cfg<D1D,Q1D,NB,NC>;
      DevPlan *pl = nullptr;
      int rc = get_plan(c, NB, &pl); if (rc) { return rc; }
      if (!pl->ucon) { set_error("mass3d_brick2: a dof has more than 8 contributions inside one batch"); return LAGB_ERR_STATE; }
      auto kern = tuned::mass3d_brick2<D1D,Q1D,NB,NC,WITH_DEN,FUSE,MINB,DBULK>;
      const size_t bytes = Cfg::smem_bytes(pl->UP, DBULK);
      rc = set_smem(c, kern, bytes); if (rc) { return rc; }
      if (WITH_DEN && pl->nbatch*NC > c.part_cap) { set_error("mass3d_brick2: partial buffer too small"); return LAGB_ERR_STATE; }
      tuned::BrickArgs a;
      a.UP = pl->UP; a.NE = c.NE; a.cstride = c.ndofs;
      a.belem = pl->belem; a.bnuniq = pl->bnuniq; a.buid = pl->buid; a.btab = pl->btab;
      a.lidx = pl->lidx; a.uoff = pl->uoff; a.upos = pl->upos; a.ucon = pl->ucon;
      a.Dq = c.d_massD; a.x = in.x; a.r = in.r; a.dold = in.dold; a.dnew = in.dnew;
      a.dinv = c.d_dinv; a.ess = c.d_essmask; a.st = c.d_state; a.comp0 = in.comp0;
      a.y = y; a.den_part = c.d_part;
      for (int col = 0; col < pl->ncolors; col++)
      {
         a.batch0 = pl->color_begin[col]; a.nbatch_launch = pl->color_begin[col + 1] - a.batch0;
         if (a.nbatch_launch <= 0) { continue; }
         cudaLaunchConfig_t cfg = {};
         cfg.gridDim = dim3((unsigned)a.nbatch_launch); cfg.blockDim = dim3(Cfg::T); cfg.dynamicSmemBytes = bytes; cfg.stream = c.stream;
         cudaLaunchAttribute at[1];
         at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
         at[0].val.programmaticStreamSerializationAllowed = 1;
         cfg.attrs = at; cfg.numAttrs = (col > 0 && c.tune[5] == 0) ? 1 : 0;
         LAGB_CUDA(cudaLaunchKernelEx(&cfg, kern, tab(c), a));
         LAGB_LAUNCH_CHECK();
      }
      if (WITH_DEN) { c.dt_nblocks = pl->nbatch; }
      return LAGB_OK;
   }
   // single-launch persistent dataflow kernel (device/mass3d_brick3.cuh); plain input only
   template<int NC, bool WITH_DEN, int NB, int MINB, bool DBULK>
   static int brick3_launch_v(Ctx &c, const MassBrickIn &in, double *y)
   {
      using Cfg = tuned::MassBrick3Cfg<D1D,Q1D,NB,NC>;
      DevPlan *pl = nullptr;
      int rc = get_plan(c, NB, &pl); if (rc) { return rc; }
      if (!pl->ucon || !pl->deps) { set_error("mass3d_brick3: schedule has no fixed-width tables"); return LAGB_ERR_STATE; }
      auto kern = tuned::mass3d_brick3<D1D,Q1D,NB,NC,WITH_DEN,MINB,DBULK>;
      const size_t bytes = Cfg::smem_bytes(pl->UP, DBULK);
      rc = set_smem(c, kern, bytes); if (rc) { return rc; }
      if (WITH_DEN && pl->nbatch*NC > c.part_cap) { set_error("mass3d_brick3: partial buffer too small"); return LAGB_ERR_STATE; }
      int &occ = c.occ_cache[(const void*)kern];
      if (occ == 0)
      {
         LAGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::T, bytes));
         if (occ < 1) { set_error("mass3d_brick3: kernel does not fit an SM"); return LAGB_ERR_STATE; }
      }
      tuned::Brick3Args g;
      tuned::BrickArgs &a = g.b;
      a.UP = pl->UP; a.NE = c.NE; a.cstride = c.ndofs; a.batch0 = 0; a.nbatch_launch = pl->nbatch;
      a.belem = pl->belem; a.bnuniq = pl->bnuniq; a.buid = pl->buid; a.btab = pl->btab;
      a.lidx = pl->lidx; a.uoff = pl->uoff; a.upos = pl->upos; a.ucon = pl->ucon;
      a.Dq = c.d_massD; a.x = in.x; a.r = nullptr; a.dold = nullptr; a.dnew = nullptr;
      a.dinv = c.d_dinv; a.ess = c.d_essmask; a.st = c.d_state; a.comp0 = in.comp0;
      a.y = y; a.den_part = c.d_part;
      g.nbatch_total = pl->nbatch; g.bmeta = reinterpret_cast<const int4*>(pl->bmeta); g.deps = pl->deps;
      g.flags = pl->flags; g.work_ctr = pl->work_ctr; g.epoch = ++pl->epoch;
      LAGB_CUDA(cudaMemsetAsync(pl->work_ctr, 0, sizeof(int), c.stream));
      const int grid = std::min(pl->nbatch, c.num_sms*occ);
      kern<<<grid, Cfg::T, bytes, c.stream>>>(tab(c), g);
      LAGB_LAUNCH_CHECK();
      if (WITH_DEN) { c.dt_nblocks = pl->nbatch; }
      return LAGB_OK;
   }
   template<int NC, bool WITH_DEN>
   static int brick3_launch(Ctx &c, const MassBrickIn &in, double *y)
   {
      if constexpr (D1D == 4 && NC == 3)   // tuning variants (lagb_tune_set key 4)
      {
         switch (c.tune[4])
         {
            case 1: return brick3_launch_v<NC,WITH_DEN,8,4,true>(c, in, y);
            case 2: return brick3_launch_v<NC,WITH_DEN,8,4,false>(c, in, y);
            case 3: return brick3_launch_v<NC,WITH_DEN,8,6,false>(c, in, y);
            case 4: return brick3_launch_v<NC,WITH_DEN,16,2,false>(c, in, y);
         }
         return brick3_launch_v<NC,WITH_DEN,8,5,false>(c, in, y);
      }
      if constexpr (D1D == 4 && NC == 1)
      {
         switch (c.tune[4])
         {
            case 1: return brick3_launch_v<NC,WITH_DEN,16,4,true>(c, in, y);
            case 2: return brick3_launch_v<NC,WITH_DEN,32,3,false>(c, in, y);
         }
         return brick3_launch_v<NC,WITH_DEN,16,6,false>(c, in, y);
      }
      return brick3_launch_v<NC,WITH_DEN,NBB,1,false>(c, in, y);
   }
   template<int NC, bool WITH_DEN, bool FUSE>
   static int brick_launch(Ctx &c, const MassBrickIn &in, double *y)
   {
      if constexpr (!FUSE) { if (c.tune[6] == 4) { return brick3_launch<NC,WITH_DEN>(c, in, y); } }
      if (c.tune[6] == 3 || c.tune[6] == 0)
      {
         if constexpr (D1D == 4 && NC == 3)   // tuning variants (lagb_tune_set key 4)
         {
            switch (c.tune[4])
            {
               case 1: return brick2_launch_v<NC,WITH_DEN,FUSE,8,4,true>(c, in, y);
               case 2: return brick2_launch_v<NC,WITH_DEN,FUSE,8,5,false>(c, in, y);
               case 3: return brick2_launch_v<NC,WITH_DEN,FUSE,16,2,true>(c, in, y);
               case 4: return brick2_launch_v<NC,WITH_DEN,FUSE,16,3,false>(c, in, y);
            }
            return brick2_launch_v<NC,WITH_DEN,FUSE,8,6,false>(c, in, y);
         }
         if constexpr (D1D == 4 && NC == 1)
         {
            switch (c.tune[4])
            {
               case 1: return brick2_launch_v<NC,WITH_DEN,FUSE,16,4,true>(c, in, y);
               case 2: return brick2_launch_v<NC,WITH_DEN,FUSE,32,3,false>(c, in, y);
               case 3: return brick2_launch_v<NC,WITH_DEN,FUSE,32,2,true>(c, in, y);
               case 4: return brick2_launch_v<NC,WITH_DEN,FUSE,8,8,false>(c, in, y);
            }
            return brick2_launch_v<NC,WITH_DEN,FUSE,16,6,false>(c, in, y);
         }
         return brick2_launch_v<NC,WITH_DEN,FUSE,NBB,1,false>(c, in, y);
      }
      if constexpr (D1D == 4)   // first brick kernel: elements per batch / resident CTAs
      {
         switch (c.tune[4])
         {
            case 1: return brick_launch_v<NC,WITH_DEN,FUSE,8,3>(c, in, y);
            case 2: return brick_launch_v<NC,WITH_DEN,FUSE,16,2>(c, in, y);
         }
         return brick_launch_v<NC,WITH_DEN,FUSE,8,4>(c, in, y);
      }
      return brick_launch_v<NC,WITH_DEN,FUSE,NBB,1>(c, in, y);
   }
   static int mass_brick(Ctx &c, int nc, const MassBrickIn &in, double *y, bool with_den)
   {
      const bool fuse = in.x == nullptr;
      if (nc == 3)
      {
         if (fuse) { return with_den ? brick_launch<3,true,true>(c, in, y) : brick_launch<3,false,true>(c, in, y); }
         return with_den ? brick_launch<3,true,false>(c, in, y) : brick_launch<3,false,false>(c, in, y);
      }
      if (nc == 1)
      {
         if (fuse) { return with_den ? brick_launch<1,true,true>(c, in, y) : brick_launch<1,false,true>(c, in, y); }
         return with_den ? brick_launch<1,true,false>(c, in, y) : brick_launch<1,false,false>(c, in, y);
      }
      set_error("mass3d_brick: nc must be 1 or 3"); return LAGB_ERR_INVALID;
   }
   static int mass_h1(Ctx &c, int nc, const double *x, double *y, bool with_den)
   {
      if (nc == 3) { return with_den ? mass_launch<3,true>(c, x, y) : mass_launch<3,false>(c, x, y); }
      if (nc == 1) { return with_den ? mass_launch<1,true>(c, x, y) : mass_launch<1,false>(c, x, y); }
      set_error("mass3d: nc must be 1 or 3"); return LAGB_ERR_INVALID;
   }
   template<int MINB, int NT = NTQ>
   static int qupdate_launch(Ctx &c, const double *S, const QPointParams &prm)
   {
      static_assert(NT % 32 == 0 && NTF % 32 == 0, "CTA sizes must be whole warps");
      using Cfg = tuned::QUpd3DCfg<D1D,Q1D>;
      auto kern = tuned::qupdate3d<D1D,Q1D,NT,MINB>;
      { int rc = set_smem(c, kern, Cfg::SMEM_BYTES); if (rc) { return rc; } }
      const int grid = c.NE;
      kern<<<grid, NT, Cfg::SMEM_BYTES, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, S, c.d_rho0DetJ0w, c.d_Jac0inv,
                                                   c.d_gamma, c.d_qweights, c.d_inv_qweights, prm, c.d_sJit, c.d_dt);
      LAGB_LAUNCH_CHECK();
      c.dt_nblocks = grid;
      return LAGB_OK;
   }
   static int qupdate(Ctx &c, const double *S, const QPointParams &prm)
   {
      if constexpr (NTQ <= 224)   // tuning variants (lagb_tune_set key 2): resident CTAs per SM
      {
         switch (c.tune[2])
         {
            case 1: return qupdate_launch<2>(c, S, prm);
            case 2: return qupdate_launch<1>(c, S, prm);
            case 3: return qupdate_launch<3>(c, S, prm);
            case 4: return qupdate_launch<5>(c, S, prm);
         }
         return qupdate_launch<4>(c, S, prm);   // 4 CTAs/SM (72 registers, L1-resident spill) beat 2 CTAs at 128: 8.1 vs 9.9 ms
      }
      else                        // Q1D >= 8: threads per element (points per thread) and resident CTAs
      {
         switch (c.tune[2])
         {
            case 1: return qupdate_launch<1,256>(c, S, prm);
            case 2: return qupdate_launch<1,512>(c, S, prm);
            case 3: return qupdate_launch<2,512>(c, S, prm);
            case 4: return qupdate_launch<2,384>(c, S, prm);
            case 5: return qupdate_launch<3,256>(c, S, prm);
         }
         if (Q1D >= 10) { return qupdate_launch<1,512>(c, S, prm); }   // 8.5 vs 11.2 ms with 256 threads (box01_hex -rs 4 -ok 5)
         return qupdate_launch<2>(c, S, prm);
      }
   }
   template<int NB, int NT, bool PF = false>
   static int force_launch(Ctx &c, const double *e, double *v)
   {
      using Cfg = tuned::Force3DCfg<D1D,Q1D>;
      auto kern = tuned::force3d<D1D,Q1D,NB,NT,PF>;
      constexpr size_t bytes = sizeof(double)*(size_t)NB*(Cfg::PER_ELEM + (PF ? 9*Cfg::NQ : 0));
      if (bytes > 227*1024) { set_error("force3d: this launch variant does not fit shared memory at this order"); return LAGB_ERR_INVALID; }
      { int rc = set_smem(c, kern, bytes); if (rc) { return rc; } }
      kern<<<(c.NE + NB - 1)/NB, NT, bytes, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, c.d_sJit, e, v);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   template<int NB, int NT, bool PF = false>
   static int forcet_launch(Ctx &c, const double *v, double *e)
   {
      using Cfg = tuned::ForceT3DCfg<D1D,Q1D>;
      auto kern = tuned::forcet3d<D1D,Q1D,NB,NT,PF>;
      constexpr size_t bytes = sizeof(double)*(size_t)NB*(Cfg::PER_ELEM + (PF ? Cfg::S_PF : 0));
      if (bytes > 227*1024) { set_error("forcet3d: this launch variant does not fit shared memory at this order"); return LAGB_ERR_INVALID; }
      { int rc = set_smem(c, kern, bytes); if (rc) { return rc; } }
      kern<<<(c.NE + NB - 1)/NB, NT, bytes, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, c.d_sJit, v, e);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   template<int NB, int NT>
   static int forcet_persist_launch(Ctx &c, const double *v, double *e)
   {
      using Cfg = tuned::ForceT3DCfg<D1D,Q1D>;
      auto kern = tuned::forcet3d_persist<D1D,Q1D,NB,NT>;
      constexpr size_t bytes = sizeof(double)*(size_t)NB*(Cfg::PER_ELEM + Cfg::S_PF);
      if (bytes > 227*1024) { set_error("forcet3d: this launch variant does not fit shared memory at this order"); return LAGB_ERR_INVALID; }
      { int rc = set_smem(c, kern, bytes); if (rc) { return rc; } }
      int &occ = c.occ_cache[(const void*)kern];
      if (occ == 0)
      {
         LAGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, bytes));
         if (occ < 1) { set_error("forcet3d_persist does not fit an SM"); return LAGB_ERR_STATE; }
      }
      const int nbatch = (c.NE + NB - 1)/NB;
      kern<<<std::min(nbatch, c.num_sms*occ), NT, bytes, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, c.d_sJit, v, e);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   static int force_mult(Ctx &c, const double *e, double *v)
   {
      if constexpr (D1D == 4)   // tuning variants (lagb_tune_set key 1)
      {
         switch (c.tune[1])
         {
            case 1: return force_launch<1,96>(c, e, v);
            case 2: return force_launch<4,256>(c, e, v);
            case 3: return force_launch<1,128>(c, e, v);
            case 4: return force_launch<1,192>(c, e, v);
         }
      }
      if constexpr (D1D >= 5)   // tuning variants of the high orders (lagb_tune_set key 1)
      {
         switch (c.tune[1])
         {
            case 1: return force_launch<NBF,NTF>(c, e, v);
            case 2: return force_launch<1,256>(c, e, v);
            case 3: return force_launch<4,256>(c, e, v);
            case 4: return force_launch<1,128,true>(c, e, v);
            case 5: return force_launch<1,256,true>(c, e, v);
         }
         // one element, 128 threads: 702 vs 1666 us (Q4Q3), 1900 vs 2569 us (Q5Q4) on box01_hex -rs 4 (profiles/r2_order_sweep.md)
         return force_launch<1,128>(c, e, v);
      }
      if constexpr (D1D == 3)
      {
         switch (c.tune[1])
         {
            case 1: return force_launch<4,128>(c, e, v);
            case 2: return force_launch<16,256>(c, e, v);
            case 3: return force_launch<8,128>(c, e, v);
            case 4: return force_launch<4,256>(c, e, v);
            case 5: return force_launch<2,64>(c, e, v);
         }
      }
      return force_launch<NBF,NTF>(c, e, v);
   }
   static int force_mult_t(Ctx &c, const double *v, double *e)
   {
      if constexpr (D1D == 4)
      {
         switch (c.tune[1])
         {
            case 1: return forcet_launch<1,64,true>(c, v, e);
            case 2: return forcet_launch<4,256>(c, v, e);
            case 3: return forcet_launch<1,128,true>(c, v, e);
            case 4: return forcet_launch<1,64>(c, v, e);
            case 5: return forcet_launch<1,96,true>(c, v, e);   // one element per CTA, cp.async prefetch of the stressJinvT slab
            case 6: return forcet_persist_launch<1,128>(c, v, e);
         }
         return forcet_persist_launch<1,96>(c, v, e);   // persistent CTAs, next element's gather in flight: 1358 vs 1452 us
      }
      if constexpr (D1D >= 5)
      {
         switch (c.tune[1])
         {
            case 1: return forcet_launch<1,128>(c, v, e);
            case 2: return forcet_launch<1,256>(c, v, e);
            case 3: return forcet_launch<4,256>(c, v, e);
            case 4: return forcet_launch<NBF,NTF>(c, v, e);
            case 5: return forcet_launch<1,(D1D == 5) ? 256 : 128,true>(c, v, e);
            case 6: return forcet_launch<1,(D1D == 5) ? 128 : 256,true>(c, v, e);
            case 7: return forcet_persist_launch<1,(D1D == 5) ? 256 : 128>(c, v, e);
         }
         // persistent CTAs, one element per batch, next element's gather in flight, cp.async stressJinvT slab:
         // Q4Q3 (128 threads) 777 us (2 elements x 256 threads: 3322, one element per CTA + slab: 1042),
         // Q5Q4 (256 threads) 2393 us (4119, 2932); key 1 = 4: the round-1 shape, 5 / 7: the other thread count
         return forcet_persist_launch<1,(D1D == 5) ? 128 : 256>(c, v, e);
      }
      if constexpr (D1D == 3)
      {
         switch (c.tune[1])
         {
            case 1: return forcet_launch<4,128>(c, v, e);
            case 2: return forcet_launch<16,256>(c, v, e);
            case 3: return forcet_launch<8,128>(c, v, e);
            case 4: return forcet_launch<4,256>(c, v, e);
            case 5: return forcet_launch<2,64>(c, v, e);
            case 6: return forcet_persist_launch<8,256>(c, v, e);
            case 7: return forcet_launch<NBF,NTF>(c, v, e);
         }
         return forcet_persist_launch<4,128>(c, v, e);   // Q2Q1: 113 us (8 elements x 256 threads per CTA: 145)
      }
      return forcet_launch<NBF,NTF>(c, v, e);
   }
   static int mass_l2(Ctx &c, const double *x, double *y)
   {
      using Cfg = tuned::MassL2Cfg<D1D,Q1D>;
      constexpr int NB = (Q1D <= 2) ? 64 : (Q1D <= 4) ? 32 : (Q1D <= 6) ? 16 : (Q1D <= 8) ? 8 : 4, NT = 256;
      auto kern = tuned::massl2_3d<D1D,Q1D,NB,NT>;
      constexpr size_t bytes = sizeof(double)*(size_t)NB*Cfg::PER_ELEM;
      { int rc = set_smem(c, kern, bytes); if (rc) { return rc; } }
      kern<<<(c.NE + NB - 1)/NB, NT, bytes, c.stream>>>(tab(c), c.NE, c.d_massD, x, y);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   static void install(KernelSet &ks)
   {
      ks.mass_h1 = &mass_h1; ks.qupdate = &qupdate; ks.force_mult = &force_mult;
      ks.force_mult_t = &force_mult_t; ks.mass_l2 = &mass_l2; ks.mass_brick = &mass_brick;
      ks.tuned_mass = true;
   }
};

bool add_tuned_kernels(KernelSet &ks, int dim, int D1D, int Q1D)
{
   if (dim != 3) { return false; }
   const int id = (D1D << 4) | Q1D;
   switch (id)
   {
      //                          D  Q  NB1 MB1 NB3 MB3 NTQ NBF  NTF
      case 0x22: TunedLaunch3D<2, 2, 64, 1, 32, 1,  32, 16, 128>::install(ks); break;
      case 0x34: TunedLaunch3D<3, 4, 32, 1, 32, 1,  64,  8, 256>::install(ks); break;
      case 0x46: TunedLaunch3D<4, 6, 32, 2, 16, 3, 224,  1,  64>::install(ks); break;
      case 0x58: TunedLaunch3D<5, 8, 16, 1,  8, 1, 256,  2, 256>::install(ks); break;
      case 0x6A: TunedLaunch3D<6, 10, 8, 1,  4, 1, 256,  1, 256>::install(ks); break;
      default: return false;
   }
   return true;
}

} // namespace lagb
