// Launchers for the tuned 3D kernels (sm_100a).
#include "ctx.hpp"
#include "device/mass3d.cuh"

namespace lagb {

template<int D1D, int Q1D, int NB1, int NB3>
struct TunedLaunch3D
{
   using Tab = DevTables<D1D,Q1D>;
   static const Tab &tab(Ctx &c) { return *reinterpret_cast<const Tab*>(c.tab_blob.data()); }

   template<int NC, bool WITH_DEN>
   static int mass_launch(Ctx &c, const double *x, double *y)
   {
      constexpr int NB = (NC == 1) ? NB1 : NB3;
      using Cfg = tuned::Mass3DCfg<D1D,Q1D,NB,NC>;
      auto kern = tuned::mass3d<D1D,Q1D,NB,NC,WITH_DEN>;
      static bool attr_set = false;
      if (!attr_set)
      {
         LAGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
         attr_set = true;
      }
      const int grid = (c.NE + NB - 1)/NB;
      if (WITH_DEN && grid*NC > c.part_cap) { set_error("mass3d: partial buffer too small"); return LAGB_ERR_STATE; }
      kern<<<grid, Cfg::T, Cfg::SMEM_BYTES, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, c.d_massD, x, y, c.d_part);
      LAGB_LAUNCH_CHECK();
      if (WITH_DEN) { c.dt_nblocks = grid; }
      return LAGB_OK;
   }
   static int mass_h1(Ctx &c, int nc, const double *x, double *y, bool with_den)
   {
      if (nc == 3) { return with_den ? mass_launch<3,true>(c, x, y) : mass_launch<3,false>(c, x, y); }
      if (nc == 1) { return with_den ? mass_launch<1,true>(c, x, y) : mass_launch<1,false>(c, x, y); }
      set_error("mass3d: nc must be 1 or 3"); return LAGB_ERR_INVALID;
   }
};

// <D1D, Q1D, NB for one component, NB for three components>: elements per CTA, chosen
// so that the shared-memory slab NC*NB*D1D*(Q1D^2+1)*8 B leaves room for several CTAs per SM.
bool add_tuned_kernels(KernelSet &ks, int dim, int D1D, int Q1D)
{
   if (dim != 3) { return false; }
   const int id = (D1D << 4) | Q1D;
   switch (id)
   {
      case 0x22: ks.mass_h1 = &TunedLaunch3D<2,2,64,32>::mass_h1; break;
      case 0x34: ks.mass_h1 = &TunedLaunch3D<3,4,32,32>::mass_h1; break;
      case 0x46: ks.mass_h1 = &TunedLaunch3D<4,6,32,16>::mass_h1; break;
      case 0x58: ks.mass_h1 = &TunedLaunch3D<5,8,16,8>::mass_h1; break;
      case 0x6A: ks.mass_h1 = &TunedLaunch3D<6,10,8,4>::mass_h1; break;
      default: return false;
   }
   ks.tuned_mass = true;
   return true;
}

} // namespace lagb
