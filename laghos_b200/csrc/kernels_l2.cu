// Launchers of the direct L2 mass solve (device/l2solve.cuh).
#include "ctx.hpp"
#include "device/l2solve.cuh"
#include <algorithm>
#include <cstdlib>

namespace lagb {

static int ensure_BL(Ctx &c);

// Builds the element inverses on first use.  They cost NE*NL^2 doubles (Q3Q2 at 64^3 elements: 1.5 GB);
// when that does not fit comfortably in free device memory the caller falls back to the CG.
static int l2_build(Ctx &c)
{
   c.l2inv_state = -1;
   const char *env = getenv("LAGB_L2_SOLVER");
   if (env && std::string(env) == "cg") { return LAGB_OK; }
   const size_t NL2 = (size_t)c.NL*c.NL, need = sizeof(double)*NL2*(size_t)c.NE;
   size_t free_b = 0, total_b = 0;
   LAGB_CUDA(cudaMemGetInfo(&free_b, &total_b));
   if (need > free_b/4) { return LAGB_OK; }
   const size_t smem = sizeof(double)*(NL2 + c.NL + (size_t)c.Q1D*c.L1D + c.NQ);
   if (smem > 200*1024) { return LAGB_OK; }
   LAGB_CUDA(cudaMalloc((void**)&c.d_l2inv, std::max<size_t>(need, 8)));
   { int rc = ensure_BL(c); if (rc) { return rc; } }
   LAGB_CUDA(cudaFuncSetAttribute(l2::l2inv_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   l2::l2inv_build<<<c.NE, 256, smem, c.stream>>>(c.dim, c.L1D, c.Q1D, c.d_BL, c.d_massD, c.d_l2inv);
   LAGB_LAUNCH_CHECK();
   c.l2inv_state = 1;
   return LAGB_OK;
}

static int ensure_BL(Ctx &c)
{
   if (c.d_BL) { return LAGB_OK; }
   LAGB_CUDA(cudaMalloc((void**)&c.d_BL, sizeof(double)*c.Q1D*std::max(1, c.L1D)));
   const double *hBL = reinterpret_cast<const double*>(c.tab_blob.data()) + 2*c.Q1D*c.D1D;
   LAGB_CUDA(cudaMemcpyAsync(c.d_BL, hBL, sizeof(double)*c.Q1D*c.L1D, cudaMemcpyHostToDevice, c.stream));
   return LAGB_OK;
}

int density_project(Ctx &c, const double *wdet, double *rho)
{
   int rc = ensure_BL(c); if (rc) { return rc; }
   const size_t smem = sizeof(double)*((size_t)c.NL*c.NL + 2*c.NL + (size_t)c.Q1D*c.L1D + 2*c.NQ);
   if (smem > 200*1024) { set_error("compute_density: element matrix does not fit shared memory"); return LAGB_ERR_INVALID; }
   LAGB_CUDA(cudaFuncSetAttribute(l2::density_project, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
   l2::density_project<<<c.NE, 256, smem, c.stream>>>(c.dim, c.L1D, c.Q1D, c.d_BL, wdet, c.d_rho0DetJ0w, rho);
   LAGB_LAUNCH_CHECK();
   return LAGB_OK;
}

template<int NLC, int EPB>
static int l2_apply_launch(Ctx &c, const double *b, double *x)
{
   constexpr int T = NLC > 0 ? ((NLC*EPB + 31)/32)*32 : 256;
   const size_t smem = sizeof(double)*(size_t)EPB*c.NL;
   auto kern = l2::l2inv_apply<NLC,EPB>;
   if (smem > 48*1024) { LAGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); }
   kern<<<(c.NE + EPB - 1)/EPB, T, smem, c.stream>>>(c.NE, c.NL, c.d_l2inv, b, x);
   LAGB_LAUNCH_CHECK();
   return LAGB_OK;
}

int l2_direct_solve(Ctx &c, const double *b, double *x, bool *done)
{
   *done = false;
   if (c.tune[10] == 1 || !c.setup_done) { return LAGB_OK; }
   if (c.l2inv_state == 0) { int rc = l2_build(c); if (rc) { return rc; } }
   if (c.l2inv_state != 1) { return LAGB_OK; }
   int rc;
   switch (c.NL)
   {
      case 1:   rc = l2_apply_launch<1,256>(c, b, x); break;     // Q1Q0
      case 4:   rc = l2_apply_launch<4,64>(c, b, x); break;      // 2D Q2Q1
      case 8:   rc = l2_apply_launch<8,32>(c, b, x); break;      // 3D Q2Q1
      case 9:   rc = l2_apply_launch<9,28>(c, b, x); break;      // 2D Q3Q2
      case 27:  rc = l2_apply_launch<27,9>(c, b, x); break;      // 3D Q3Q2
      case 64:  rc = l2_apply_launch<64,4>(c, b, x); break;      // 3D Q4Q3
      default:  rc = l2_apply_launch<0,2>(c, b, x); break;
   }
   if (rc) { return rc; }
   *done = true;
   return LAGB_OK;
}

} // namespace lagb
