// Launchers for the generic one-thread-per-element kernels (all reference kernel
// ids, reference laghos_assembly.cpp:536-548 / laghos_solver.cpp:1387-1396, + 3D (6,10)).
#include "ctx.hpp"
#include "device/generic_kernels.cuh"

namespace lagb {

template<int DIM, int D1D, int Q1D>
struct GenericLaunch
{
   using Tab = DevTables<D1D,Q1D>;
   static constexpr int BS = 64;
   static const Tab &tab(Ctx &c) { return *reinterpret_cast<const Tab*>(c.tab_blob.data()); }
   static int grid(const Ctx &c) { return (c.NE + BS - 1)/BS; }

   static int mass_h1(Ctx &c, int nc, const double *x, double *y, bool with_den)
   {
      for (int k = 0; k < nc; k++)
      {
         generic::mass_h1<DIM,D1D,Q1D><<<grid(c), BS, 0, c.stream>>>(tab(c), c.NE, c.d_map, c.d_massD,
                                                                  x + k*c.ndofs, y + k*c.ndofs);
         LAGB_LAUNCH_CHECK();
      }
      (void)with_den;
      return LAGB_OK;
   }
   static int mass_diag(Ctx &c, double *diag)
   {
      generic::mass_h1_diag<DIM,D1D,Q1D><<<grid(c), BS, 0, c.stream>>>(tab(c), c.NE, c.d_map, c.d_massD, diag);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   static int mass_l2(Ctx &c, const double *x, double *y)
   {
      generic::mass_l2<DIM,D1D,Q1D><<<grid(c), BS, 0, c.stream>>>(tab(c), c.NE, c.d_massD, x, y);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   static int force_mult(Ctx &c, const double *e, double *v)
   {
      generic::force_mult<DIM,D1D,Q1D><<<grid(c), BS, 0, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, c.d_sJit, e, v);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   static int force_mult_t(Ctx &c, const double *v, double *e)
   {
      generic::force_mult_t<DIM,D1D,Q1D><<<grid(c), BS, 0, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, c.d_sJit, v, e);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   static int qupdate(Ctx &c, const double *S, const QPointParams &prm)
   {
      const int g = grid(c);
      generic::qupdate<DIM,D1D,Q1D><<<g, BS, 0, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, S, c.d_rho0DetJ0w,
                                                           c.d_Jac0inv, c.d_gamma, c.d_qweights, prm, c.d_sJit, c.d_dt);
      LAGB_LAUNCH_CHECK();
      c.dt_nblocks = g;
      return LAGB_OK;
   }
   static int rho0detj0(Ctx &c, const double *x0, const double *rho0_gf, const double *rho0_q, double *elem_vol)
   {
      generic::rho0detj0<DIM,D1D,Q1D><<<grid(c), BS, 0, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, x0, rho0_gf, rho0_q,
                                                                   c.d_qweights, c.d_rho0DetJ0w, c.d_Jac0inv, c.d_massD, elem_vol);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   static int taylor(Ctx &c, const double *x, double *esrc)
   {
      generic::taylor_source<DIM,D1D,Q1D><<<grid(c), BS, 0, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, x, c.d_qweights, esrc);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   static int detj_w(Ctx &c, const double *x, double *out)
   {
      generic::detj_w<DIM,D1D,Q1D><<<grid(c), BS, 0, c.stream>>>(tab(c), c.NE, c.ndofs, c.d_map, x, c.d_qweights, out);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   static KernelSet make()
   {
      KernelSet k;
      k.mass_h1 = &mass_h1; k.mass_diag = &mass_diag; k.mass_l2 = &mass_l2;
      k.force_mult = &force_mult; k.force_mult_t = &force_mult_t; k.qupdate = &qupdate;
      k.rho0detj0 = &rho0detj0; k.taylor = &taylor; k.detj_w = &detj_w;
      return k;
   }
};

KernelSet make_generic_kernels(int dim, int D1D, int Q1D)
{
   const int id = (dim << 8) | (D1D << 4) | Q1D;
   switch (id)
   {
      case 0x222: return GenericLaunch<2,2,2>::make();
      case 0x234: return GenericLaunch<2,3,4>::make();
      case 0x246: return GenericLaunch<2,4,6>::make();
      case 0x258: return GenericLaunch<2,5,8>::make();
      case 0x26A: return GenericLaunch<2,6,10>::make();
      case 0x322: return GenericLaunch<3,2,2>::make();
      case 0x334: return GenericLaunch<3,3,4>::make();
      case 0x346: return GenericLaunch<3,4,6>::make();
      case 0x358: return GenericLaunch<3,5,8>::make();
      case 0x36A: return GenericLaunch<3,6,10>::make();
   }
   return KernelSet();
}

} // namespace lagb
