// C ABI implementation (include/laghos_b200.h): context, operator entry points,
// device-resident PCG driver, timers, NCCL plumbing.
#include "ctx.hpp"
#include "host/batch_plan.hpp"
#include <algorithm>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <mutex>

namespace lagb {

static thread_local std::string g_err;
int64_t g_launch_count = 0;
void set_error(const std::string &msg) { g_err = msg; }

// CTAs of a PCG finish kernel: a small grid with a last-arrival second stage pays off above a few thousand partials
static int fin_grid(int npart) { return npart > 4096 ? pcg::FIN_CTAS : 1; }

int vec_grid(int64_t n)
{
   const int64_t b = (n + pcg::RB - 1)/pcg::RB;
   return (int)std::max<int64_t>(1, std::min<int64_t>(b, 148*16));
}

// One-wave grid of a grid-stride vector kernel: resident CTAs per SM (occupancy API, cached per context) x SMs.
// vec_grid() assumes 16 CTAs per SM; the load-first PCG kernels hold 80-114 registers (2-3 CTAs per SM), so that grid
// ran in 4-8 waves of CTAs with 3 loop iterations each and a block reduction per CTA (update_r: 60 % of the HBM peak).
template<typename K>
static int wave_grid(Ctx &c, K kern, int64_t n)
{
   int &occ = c.occ_cache[(const void*)kern];
   if (occ == 0)
   {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, pcg::RB, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 2; }
   }
   const int64_t b = (n + pcg::RB - 1)/pcg::RB;
   return (int)std::max<int64_t>(1, std::min<int64_t>(b, (int64_t)c.num_sms*occ));
}

// ---- timers: CUDA events on the context stream, resolved lazily ----
static cudaEvent_t timer_event(Ctx &c)
{
   if (!c.timer.pool.empty()) { cudaEvent_t e = c.timer.pool.back(); c.timer.pool.pop_back(); return e; }
   cudaEvent_t e; cudaEventCreate(&e); return e;
}
static int timer_resolve(Ctx &c)
{
   for (int w = 0; w < Timer::NT; w++)
   {
      auto &pend = c.timer.pending[w];
      // an interval whose end event is not recorded yet (timers nest) stays pending
      const bool keep_last = !pend.empty() && c.timer.open[w];
      const size_t n = pend.size() - (keep_last ? 1 : 0);
      for (size_t i = 0; i < n; i++)
      {
         auto &p = pend[i];
         LAGB_CUDA(cudaEventSynchronize(p.second));
         float ms = 0.f; LAGB_CUDA(cudaEventElapsedTime(&ms, p.first, p.second));
         c.timer.acc[w] += 1e-3*ms;
         c.timer.pool.push_back(p.first); c.timer.pool.push_back(p.second);
      }
      pend.erase(pend.begin(), pend.begin() + n);
   }
   return LAGB_OK;
}
int timer_begin(Ctx &c, int w)
{
   if (c.timer.open[w] && !c.timer.pending[w].empty())
   {
      // an error return left this interval without its end event: drop it
      auto p = c.timer.pending[w].back(); c.timer.pending[w].pop_back();
      c.timer.pool.push_back(p.first); c.timer.pool.push_back(p.second);
      c.timer.open[w] = false;
   }
   cudaEvent_t a = timer_event(c), b = timer_event(c);
   LAGB_CUDA(cudaEventRecord(a, c.stream));
   c.timer.pending[w].push_back({a, b});
   c.timer.open[w] = true;
   return LAGB_OK;
}
int timer_end(Ctx &c, int w)
{
   LAGB_CUDA(cudaEventRecord(c.timer.pending[w].back().second, c.stream));
   c.timer.open[w] = false;
   if (c.timer.pending[w].size() > 4096) { return timer_resolve(c); }
   return LAGB_OK;
}

// ---- NCCL through dlopen (torch ships libnccl.so.2; no link-time dependency) ----
struct NcclApi
{
   void *lib = nullptr;
   int (*GetUniqueId)(void *id) = nullptr;
   int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
   int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
   int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
   int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
   int (*GroupStart)() = nullptr;
   int (*GroupEnd)() = nullptr;
   int (*CommDestroy)(void*) = nullptr;
   const char *(*GetErrorString)(int) = nullptr;
   void *CommInitRankRaw = nullptr;
};
struct Uid128 { char internal[128]; };
static NcclApi g_nccl;
static int nccl_load()
{
   if (g_nccl.lib) { return LAGB_OK; }
   const char *names[] = {"libnccl.so.2", "libnccl.so"};
   for (const char *n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) { break; } }
   if (!g_nccl.lib) { set_error("cannot dlopen libnccl.so.2"); return LAGB_ERR_NCCL; }
   auto sym = [&](const char *s) { return dlsym(g_nccl.lib, s); };
   g_nccl.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
   g_nccl.CommInitRankRaw = sym("ncclCommInitRank");
   g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
   g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
   g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
   g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
   g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
   g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
   g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
   g_nccl.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
   if (!g_nccl.GetUniqueId || !g_nccl.CommInitRankRaw || !g_nccl.AllReduce || !g_nccl.Send || !g_nccl.Recv)
   { set_error("libnccl: missing symbols"); return LAGB_ERR_NCCL; }
   return LAGB_OK;
}
#define LAGB_NCCL(call) do { int r__ = (call); if (r__ != 0) { \
   lagb::set_error(std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error")); return LAGB_ERR_NCCL; } } while (0)
// nccl enums: ncclFloat64 = 8, ncclSum = 0, ncclMin = 3 (stable across NCCL 2.x)
static const int NCCL_F64 = 8, NCCL_SUM = 0, NCCL_MIN = 3;

int allreduce_sum(Ctx &c, double *d_vals, int n)
{
   if (c.nranks <= 1) { return LAGB_OK; }
   LAGB_NCCL(g_nccl.AllReduce(d_vals, d_vals, n, NCCL_F64, NCCL_SUM, c.nccl_comm, c.stream));
   return LAGB_OK;
}
int allreduce_min(Ctx &c, double *d_vals, int n)
{
   if (c.nranks <= 1) { return LAGB_OK; }
   LAGB_NCCL(g_nccl.AllReduce(d_vals, d_vals, n, NCCL_F64, NCCL_MIN, c.nccl_comm, c.stream));
   return LAGB_OK;
}

__global__ void halo_pack(int n, int nc, int64_t cstride, const int *__restrict__ idx,
                          const double *__restrict__ v, double *__restrict__ buf)
{
   for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < n*nc; i += gridDim.x*blockDim.x)
   {
      const int c = i / n, k = i - c*n;
      buf[i] = v[idx[k] + c*cstride];
   }
}
__global__ void halo_add(int n, int nc, int64_t cstride, const int *__restrict__ idx,
                         const double *__restrict__ buf, double *__restrict__ v)
{
   for (int i = blockIdx.x*blockDim.x + threadIdx.x; i < n*nc; i += gridDim.x*blockDim.x)
   {
      const int c = i / n, k = i - c*n;
      v[idx[k] + c*cstride] += buf[i];
   }
}

// single-phase exchange: pack every (neighbour, shared dof) entry, message layout per
// neighbour k: [c][j], base offset nc*off[k]
__global__ void halo_pack_all(int total, int nc, int64_t cstride, const int *__restrict__ idx,
                              const unsigned char *__restrict__ nbk, const int *__restrict__ off, const int *__restrict__ cnt,
                              const double *__restrict__ v, double *__restrict__ buf)
{
   for (int J = blockIdx.x*blockDim.x + threadIdx.x; J < total; J += gridDim.x*blockDim.x)
   {
      const int k = nbk[J], o = off[k], n = cnt[k], j = J - o;
      const int id = idx[J];
      for (int c = 0; c < nc; c++) { buf[(size_t)nc*o + (size_t)c*n + j] = v[id + c*cstride]; }
   }
}
// v[dof] = sum over the sharers of dof (own value and received values) in ascending rank order
__global__ void halo_combine(int nu, int nc, int64_t cstride, const int *__restrict__ u_dof, const int *__restrict__ u_ptr,
                             const int *__restrict__ u_src, const unsigned char *__restrict__ nbk,
                             const int *__restrict__ off, const int *__restrict__ cnt,
                             const double *__restrict__ recv, double *__restrict__ v)
{
   for (int u = blockIdx.x*blockDim.x + threadIdx.x; u < nu; u += gridDim.x*blockDim.x)
   {
      const int dof = u_dof[u], p0 = u_ptr[u], p1 = u_ptr[u + 1];
      for (int c = 0; c < nc; c++)
      {
         const double own = v[dof + c*cstride];
         double acc = 0.0;
         for (int p = p0; p < p1; p++)
         {
            const int J = u_src[p];
            if (J < 0) { acc += own; }
            else
            {
               const int k = nbk[J], o = off[k], n = cnt[k];
               acc += recv[(size_t)nc*o + (size_t)c*n + (J - o)];
            }
         }
         v[dof + c*cstride] = acc;
      }
   }
}

// Sum the partial values of shared dofs over the ranks that share them
// (P^t then P in the reference's RAP operator, laghos_assembly.cpp:95 / SURVEY 8e):
// per phase, pack -> grouped ncclSend/ncclRecv -> add.  After all phases every
// sharing rank holds the same fully summed value.
// peer-memory exchange in two halves so that work can be enqueued in between: pack straight into the neighbours'
// receive areas and raise the flags / wait for the neighbours' flags and combine
static bool halo_p2p(const Ctx &c) { return c.nranks > 1 && !c.nbrs.empty() && c.halo_single && c.p2p_on && c.tune[11] == 0; }
int halo_begin_p2p(Ctx &c, const double *v, int nc)
{
   const unsigned long long seq = ++c.p2p_halo_seq;
   const int nnbr = (int)c.nbrs.size();
   const int g = std::max(1, std::min(296, (c.halo_total + 255)/256));
   LAGB_LAUNCH_K(c, p2p::halo_pack_p2p, g, 256, 0, c.p2p_dev, seq, c.halo_total, nc, (int64_t)c.ndofs, (const int*)c.d_pack_idx,
                 (const unsigned char*)c.d_pack_nb, (const int*)c.d_nbr_off, (const int*)c.d_nbr_n, (const int*)c.d_nbr_roff,
                 (const int*)c.d_nbr_rank, nnbr, v, c.d_pack_done);
   return LAGB_OK;
}
int halo_end_p2p(Ctx &c, double *v, int nc)
{
   const unsigned long long seq = c.p2p_halo_seq;
   const int nnbr = (int)c.nbrs.size();
   LAGB_LAUNCH_K(c, p2p::halo_combine_p2p, std::max(1, std::min(296, (c.halo_nu + 255)/256)), 256, 0, c.p2p_dev, seq, c.halo_nu, nc,
                 (int64_t)c.ndofs, (const int*)c.d_u_dof, (const int*)c.d_u_ptr, (const int*)c.d_u_src, (const unsigned char*)c.d_pack_nb,
                 (const int*)c.d_nbr_off, (const int*)c.d_nbr_n, (const int*)c.d_nbr_rank, nnbr, v);
   return LAGB_OK;
}

int halo_sum(Ctx &c, double *v, int nc)
{
   if (c.nranks <= 1 || c.nbrs.empty()) { return LAGB_OK; }
   if (halo_p2p(c))
   {
      int rc = halo_begin_p2p(c, v, nc); if (rc) { return rc; }
      return halo_end_p2p(c, v, nc);
   }
   if (c.halo_single)
   {
      const int g = std::max(1, std::min(1024, (c.halo_total + 255)/256));
      halo_pack_all<<<g, 256, 0, c.stream>>>(c.halo_total, nc, c.ndofs, c.d_pack_idx, c.d_pack_nb, c.d_nbr_off, c.d_nbr_n, v, c.d_send_all);
      LAGB_LAUNCH_CHECK();
      LAGB_NCCL(g_nccl.GroupStart());
      for (size_t k = 0; k < c.nbrs.size(); k++)
      {
         const size_t o = (size_t)nc*c.h_nbr_off[k], n = (size_t)nc*c.h_nbr_n[k];
         LAGB_NCCL(g_nccl.Send(c.d_send_all + o, n, NCCL_F64, c.nbrs[k].rank, c.nccl_comm, c.stream));
         LAGB_NCCL(g_nccl.Recv(c.d_recv_all + o, n, NCCL_F64, c.nbrs[k].rank, c.nccl_comm, c.stream));
      }
      LAGB_NCCL(g_nccl.GroupEnd());
      halo_combine<<<std::max(1, std::min(1024, (c.halo_nu + 255)/256)), 256, 0, c.stream>>>(
         c.halo_nu, nc, c.ndofs, c.d_u_dof, c.d_u_ptr, c.d_u_src, c.d_pack_nb, c.d_nbr_off, c.d_nbr_n, c.d_recv_all, v);
      LAGB_LAUNCH_CHECK();
      return LAGB_OK;
   }
   for (int ph = 0; ph < c.nphases; ph++)
   {
      bool any = false;
      for (auto &nb : c.nbrs)
      {
         if (nb.phase != ph || nb.n == 0) { continue; }
         any = true;
         halo_pack<<<std::max(1, std::min(1024, (nb.n*nc + 255)/256)), 256, 0, c.stream>>>(nb.n, nc, c.ndofs, nb.d_idx, v, nb.d_send);
         LAGB_LAUNCH_CHECK();
      }
      if (!any) { continue; }
      LAGB_NCCL(g_nccl.GroupStart());
      for (auto &nb : c.nbrs)
      {
         if (nb.phase != ph || nb.n == 0) { continue; }
         LAGB_NCCL(g_nccl.Send(nb.d_send, (size_t)nb.n*nc, NCCL_F64, nb.rank, c.nccl_comm, c.stream));
         LAGB_NCCL(g_nccl.Recv(nb.d_recv, (size_t)nb.n*nc, NCCL_F64, nb.rank, c.nccl_comm, c.stream));
      }
      LAGB_NCCL(g_nccl.GroupEnd());
      for (auto &nb : c.nbrs)
      {
         if (nb.phase != ph || nb.n == 0) { continue; }
         halo_add<<<std::max(1, std::min(1024, (nb.n*nc + 255)/256)), 256, 0, c.stream>>>(nb.n, nc, c.ndofs, nb.d_idx, nb.d_recv, v);
         LAGB_LAUNCH_CHECK();
      }
   }
   return LAGB_OK;
}

__global__ void build_dinv(int64_t n, const double *__restrict__ diag, double *__restrict__ dinv)
{
   for (int64_t i = blockIdx.x*(int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x)
   {
      dinv[i] = 1.0/diag[i];
   }
}
__global__ void neg_inplace(double *y, int64_t n)
{
   for (int64_t i = blockIdx.x*(int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x*blockDim.x) { y[i] = -y[i]; }
}

template<typename T> static int dev_alloc(T **p, size_t n)
{
   LAGB_CUDA(cudaMalloc((void**)p, std::max<size_t>(n, 1)*sizeof(T)));
   return LAGB_OK;
}
template<typename T> static int dev_upload(T **p, const T *h, size_t n)
{
   int r = dev_alloc(p, n); if (r) { return r; }
   if (n) { LAGB_CUDA(cudaMemcpy(*p, h, n*sizeof(T), cudaMemcpyHostToDevice)); }
   return LAGB_OK;
}

// Brick schedule of the H1 mass apply for NB elements per batch: built on the host from the gather
// map on first use, uploaded once (host/batch_plan.hpp).
int get_plan(Ctx &c, int NB, DevPlan **out)
{
   const int shape_sel = c.tune[7];
   const int key = NB | (shape_sel << 16);
   auto it = c.plans.find(key);
   if (it != c.plans.end()) { *out = &it->second; return LAGB_OK; }
   BatchPlan bp; std::string err;
   int shape[3] = {0, 0, 0};
   if (shape_sel > 0 && c.elem_grid[0] > 0)
   {
      // x-long bricks: rows of (D1D-1)*bx + 1 contiguous L-vector entries per lattice row
      int bx = (shape_sel == 2) ? NB : std::min(NB, 4), rem = NB/bx, by = std::min(rem, 2), bz = rem/by;
      while (bx > c.elem_grid[0] && bx > 1) { bx /= 2; }
      shape[0] = bx; shape[1] = by; shape[2] = bz;
      if (by > c.elem_grid[1] || bz > c.elem_grid[2] || bx*by*bz > NB) { shape[0] = shape[1] = shape[2] = 0; }
   }
   if (bp.build(c.h_map.data(), c.NE, c.ND, c.ndofs, c.elem_grid, NB, shape, err)) { set_error(err); return LAGB_ERR_INVALID; }
   DevPlan dp;
   dp.NB = NB; dp.UP = bp.UP; dp.nbatch = bp.nbatch; dp.ncolors = bp.ncolors; dp.ntab = bp.ntab; dp.color_begin = bp.color_begin;
   for (int k = 0; k < 3; k++) { dp.brick[k] = bp.brick[k]; }
   // lidx rows padded to a multiple of 8 entries (16-byte vector loads)
   const int NDP = ((c.ND + 7)/8)*8;
   std::vector<uint16_t> lp((size_t)bp.ntab*NB*NDP, 0);
   for (int t = 0; t < bp.ntab; t++)
      for (int e = 0; e < NB; e++)
         for (int i = 0; i < c.ND; i++) { lp[((size_t)t*NB + e)*NDP + i] = bp.lidx[((size_t)t*NB + e)*c.ND + i]; }
   // fixed-width contribution table of the second brick kernel: plane slot (e*D1D + dz)*PLANE + dxy
   // of every element-local dof that maps to the unique slot (3D only)
   std::vector<uint16_t> uc;
   if (c.dim == 3)
   {
      const int DD = c.D1D*c.D1D, QQ = c.Q1D*c.Q1D, PLANE = (std::max(QQ, DD) | 1);
      uc.assign((size_t)bp.ntab*bp.UP*8, (uint16_t)0xffff);
      bool ok = (size_t)NB*c.D1D*PLANE < 0xffff;
      for (int t = 0; t < bp.ntab && ok; t++)
         for (int s = 0; s < bp.UP && ok; s++)
         {
            const int p0 = bp.uoff[(size_t)t*(bp.UP + 1) + s], p1 = bp.uoff[(size_t)t*(bp.UP + 1) + s + 1];
            if (p1 - p0 > 8) { ok = false; break; }
            for (int p = p0; p < p1; p++)
            {
               const int pos = bp.upos[(size_t)t*NB*c.ND + p], e = pos / c.ND, i = pos % c.ND;
               uc[((size_t)t*bp.UP + s)*8 + (p - p0)] = (uint16_t)((e*c.D1D + i / DD)*PLANE + i % DD);
            }
         }
      if (!ok) { uc.clear(); }
   }
   int rc = 0;
   rc |= dev_upload(&dp.belem, bp.elem.data(), bp.elem.size());
   rc |= dev_upload(&dp.bnuniq, bp.nuniq.data(), bp.nuniq.size());
   rc |= dev_upload(&dp.btab, bp.tab.data(), bp.tab.size());
   rc |= dev_upload(&dp.buid, bp.uid.data(), bp.uid.size());
   rc |= dev_upload(&dp.lidx, lp.data(), lp.size());
   rc |= dev_upload(&dp.uoff, bp.uoff.data(), bp.uoff.size());
   rc |= dev_upload(&dp.upos, bp.upos.data(), bp.upos.size());
   if (!uc.empty()) { rc |= dev_upload(&dp.ucon, uc.data(), uc.size()); }
   if (bp.deps_ok)
   {
      std::vector<int> bm((size_t)bp.nbatch*4, 0);
      for (int k = 0; k < bp.nbatch; k++)
      {
         int nel = 0; for (int j = 0; j < NB; j++) { nel += bp.elem[(size_t)k*NB + j] >= 0; }
         bm[4*(size_t)k] = nel; bm[4*(size_t)k + 1] = bp.nuniq[k]; bm[4*(size_t)k + 2] = bp.tab[k];
      }
      rc |= dev_upload(&dp.bmeta, bm.data(), bm.size());
      rc |= dev_upload(&dp.deps, bp.deps.data(), bp.deps.size());
      rc |= dev_alloc(&dp.flags, (size_t)bp.nbatch);
      rc |= dev_alloc(&dp.work_ctr, 1);
      if (!rc) { LAGB_CUDA(cudaMemset(dp.flags, 0, sizeof(int)*(size_t)bp.nbatch)); }
   }
   if (rc) { return LAGB_ERR_CUDA; }
   auto ins = c.plans.emplace(key, dp);
   *out = &ins.first->second;
   return LAGB_OK;
}

// Peer-memory setup (device/p2p.cuh): one communication buffer per rank, mapped by every other rank through
// cudaIpc handles that are exchanged with one ncclAllGather; the table "where does my message start in your
// receive area" comes from a second all-gather of the neighbour offsets.  Any failure (no peer access, IPC not
// permitted, a neighbour rank listed twice) leaves p2p_on false on ALL ranks and the NCCL path in use.
static int p2p_setup(Ctx &c, int nnbr, const int32_t *nbr_rank)
{
   c.p2p_on = false;
   const char *env = getenv("LAGB_P2P");
   if ((env && std::string(env) == "0") || c.nranks < 2 || c.nranks > p2p::MAXR || !g_nccl.AllGather) { return LAGB_OK; }
   const int R = c.nranks;
   // halo capacity: the largest concatenated shared-entry count over the ranks
   double *d_x = c.d_tmp + 12;
   {
      double hv = (double)(c.halo_single ? c.halo_total : 0);
      LAGB_CUDA(cudaMemcpyAsync(d_x, &hv, sizeof(double), cudaMemcpyHostToDevice, c.stream));
      LAGB_NCCL(g_nccl.AllReduce(d_x, d_x, 1, NCCL_F64, 2 /* ncclMax */, c.nccl_comm, c.stream));
      LAGB_CUDA(cudaMemcpyAsync(&hv, d_x, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      LAGB_CUDA(cudaStreamSynchronize(c.stream));
      const size_t cap = (size_t)hv;
      p2p::Layout &L = c.p2p_dev.lay;
      size_t o = 0;
      auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255)/256*256; return at; };
      L.scal = take(sizeof(double)*2*R*p2p::SLOTW);
      L.sflag = take(sizeof(unsigned long long)*2*R);
      L.hflag = take(sizeof(unsigned long long)*2*R);
      L.err = take(sizeof(int));
      L.halo[0] = take(sizeof(double)*3*cap);
      L.halo[1] = take(sizeof(double)*3*cap);
      L.bytes = o;
   }
   bool ok = true;
   if (cudaMalloc((void**)&c.p2p_base, c.p2p_dev.lay.bytes) != cudaSuccess) { cudaGetLastError(); ok = false; c.p2p_base = nullptr; }
   if (ok) { LAGB_CUDA(cudaMemset(c.p2p_base, 0, c.p2p_dev.lay.bytes)); }
   // exchange handles (64 bytes each) and the neighbour offset rows (R ints each)
   struct Msg { cudaIpcMemHandle_t h; int ok; int off[p2p::MAXR]; };
   Msg mine; memset(&mine, 0, sizeof mine);
   if (ok && cudaIpcGetMemHandle(&mine.h, c.p2p_base) != cudaSuccess) { cudaGetLastError(); ok = false; }
   for (int r = 0; r < R; r++) { mine.off[r] = -1; }
   for (int k = 0; k < nnbr && c.halo_single; k++)
   {
      if (nbr_rank[k] < 0 || nbr_rank[k] >= R || mine.off[nbr_rank[k]] >= 0) { ok = false; break; }   // listed twice
      mine.off[nbr_rank[k]] = c.h_nbr_off[k];
   }
   mine.ok = ok ? 1 : 0;
   Msg *d_msgs = nullptr;
   LAGB_CUDA(cudaMalloc((void**)&d_msgs, sizeof(Msg)*(R + 1)));
   LAGB_CUDA(cudaMemcpyAsync(d_msgs + R, &mine, sizeof(Msg), cudaMemcpyHostToDevice, c.stream));
   LAGB_NCCL(g_nccl.AllGather(d_msgs + R, d_msgs, sizeof(Msg), 0 /* ncclInt8 */, c.nccl_comm, c.stream));
   std::vector<Msg> all(R);
   LAGB_CUDA(cudaMemcpyAsync(all.data(), d_msgs, sizeof(Msg)*R, cudaMemcpyDeviceToHost, c.stream));
   LAGB_CUDA(cudaStreamSynchronize(c.stream));
   cudaFree(d_msgs);
   for (int r = 0; r < R; r++) { ok = ok && all[r].ok; }
   // map the peers (every rank must have opened every handle before anyone frees: buffers live as long as the context)
   for (int r = 0; r < R; r++) { c.p2p_dev.peer[r] = nullptr; }
   for (int r = 0; r < R && ok; r++)
   {
      if (r == c.rank) { c.p2p_dev.peer[r] = c.p2p_base; continue; }
      void *p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
      c.p2p_opened.push_back(p);
      c.p2p_dev.peer[r] = (char*)p;
   }
   // collective decision: everyone or no one
   {
      double hv = ok ? 1.0 : 0.0;
      LAGB_CUDA(cudaMemcpyAsync(d_x, &hv, sizeof(double), cudaMemcpyHostToDevice, c.stream));
      LAGB_NCCL(g_nccl.AllReduce(d_x, d_x, 1, NCCL_F64, NCCL_MIN, c.nccl_comm, c.stream));
      LAGB_CUDA(cudaMemcpyAsync(&hv, d_x, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      LAGB_CUDA(cudaStreamSynchronize(c.stream));
      ok = hv > 0.5;
   }
   if (!ok) { return LAGB_OK; }
   c.p2p_dev.rank = c.rank; c.p2p_dev.nranks = R;
   if (c.halo_single)
   {
      std::vector<int> roff(nnbr), nr(nnbr);
      for (int k = 0; k < nnbr; k++)
      {
         nr[k] = nbr_rank[k];
         roff[k] = all[nbr_rank[k]].off[c.rank];      // my message's offset in that rank's receive area
         if (roff[k] < 0) { return LAGB_OK; }          // asymmetric neighbour lists: stay on NCCL (decided before any use)
      }
      int rc = dev_upload(&c.d_nbr_roff, roff.data(), roff.size()); if (rc) { return rc; }
      rc = dev_upload(&c.d_nbr_rank, nr.data(), nr.size()); if (rc) { return rc; }
   }
   { int rc = dev_alloc(&c.d_pack_done, 1); if (rc) { return rc; } LAGB_CUDA(cudaMemset(c.d_pack_done, 0, sizeof(unsigned int))); }
   { int rc = dev_upload(&c.d_p2p_dev, &c.p2p_dev, 1); if (rc) { return rc; } }
   c.p2p_on = true;
   return LAGB_OK;
}

static int ipow(int a, int b) { int r = 1; while (b-- > 0) { r *= a; } return r; }

// ---------------------------------------------------------------------------
// PCG driver (MFEM CGSolver::Mult) on device-resident scalars
// ---------------------------------------------------------------------------
template<int NC>
static int pcg_run_nc(Ctx &c, bool l2, int comp0, const double *b, double *x, double rel_tol, int max_iter,
                      bool iterative_mode, int *h_iters)
{
   const int64_t n = l2 ? c.ndofs_l2 : c.ndofs;
   const int64_t cs = n;
   double *r = l2 ? c.d_lr : c.d_r, *d = l2 ? c.d_ld : c.d_d, *z = l2 ? c.d_lz : c.d_z;
   pcg::Prec P;
   P.dinv = l2 ? nullptr : c.d_dinv; P.ess = l2 ? nullptr : c.d_essmask; P.comp0 = comp0;
   const unsigned char *own = l2 ? nullptr : c.d_own;
   const int g = vec_grid(n);
   if (g*NC > c.part_cap) { set_error("pcg: partial buffer too small"); return LAGB_ERR_STATE; }
   // one-wave grids of the two per-iteration vector kernels (lagb_tune_set key 15 = 1: the fixed 148*16 grid)
   const bool wave = c.tune[15] == 0;
   const int g_r = wave ? wave_grid(c, pcg::update_r<NC>, n) : g;
   const int g_dx = wave ? std::min(wave_grid(c, pcg::update_dx<NC,true>, n), wave_grid(c, pcg::update_dx<NC,false>, n)) : g;
   KernelSet &ks = (c.variant == 1) ? c.ks_generic : c.ks;
   // H1 apply through the atomic-free brick kernels (plain stores): no zero fill of z anywhere
   const bool bapply = !l2 && c.variant == 0 && ks.mass_brick != nullptr && c.tune[6] >= 2 && (NC == 1 || NC == 3);
   const bool zero_z = !l2 && !bapply;   // only the atomic scatter accumulates into z
   // Zero fill of A d off the critical path: two result buffers alternate, and the one an iteration has just
   // consumed (update_r) is cleared on an auxiliary stream while the direction update and the NEXT operator apply
   // run (the mass kernel is LSU / fp64 bound at 36 % of the HBM bandwidth, so the 172 MB memset is free there).
   // Invariant: a buffer not in use is zero, or its memset is pending behind c.ev_zclr[i].
   const bool zalt = zero_z && c.tune[13] == 0 && c.aux_stream != nullptr && c.tune[8] == 0;
   double *zb[2] = {c.d_z, c.d_d2};
   int zi = 0;
   auto z_acquire = [&](int i) -> int
   {
      if (c.z_pending[i]) { LAGB_CUDA(cudaStreamWaitEvent(c.stream, c.ev_zclr[i], 0)); c.z_pending[i] = false; }
      return LAGB_OK;
   };
   auto z_release = [&](int i) -> int     // after the last reader of zb[i] has been enqueued
   {
      LAGB_CUDA(cudaEventRecord(c.ev_zuse, c.stream));
      LAGB_CUDA(cudaStreamWaitEvent(c.aux_stream, c.ev_zuse, 0));
      LAGB_CUDA(cudaMemsetAsync(zb[i], 0, sizeof(double)*NC*n, c.aux_stream));
      LAGB_CUDA(cudaEventRecord(c.ev_zclr[i], c.aux_stream));
      c.z_pending[i] = true;
      return LAGB_OK;
   };
   if (zalt)
   {
      if (!c.z_clean)
      {
         // another path (brick kernels, experiments) used the buffers as scratch: restore the invariant once
         LAGB_CUDA(cudaStreamSynchronize(c.aux_stream));
         c.z_pending[0] = c.z_pending[1] = false;
         LAGB_CUDA(cudaMemsetAsync(zb[0], 0, sizeof(double)*(size_t)c.ndofs*c.dim, c.stream));
         LAGB_CUDA(cudaMemsetAsync(zb[1], 0, sizeof(double)*(size_t)c.ndofs*c.dim, c.stream));
         c.z_clean = true;
      }
      int rz = z_acquire(0); if (rz) { return rz; }
      z = zb[0];
   }

   // apply: z (+)= A v ; returns number of partial blocks if the kernel produced d^t A d partials
   size_t den_src_off = 0;     // where the operator's d^t A d partials start inside d_part
   auto apply = [&](const double *v, bool want_den, int &den_blocks) -> int
   {
      den_blocks = 0; den_src_off = 0;
      if (l2) { return ks.mass_l2(c, v, z); }
      if (c.profile_mass) { int rt = timer_begin(c, 4); if (rt) { return rt; } }
      int rc;
      if (bapply) { MassBrickIn in; in.x = v; rc = ks.mass_brick(c, NC, in, z, want_den); }
      else { rc = ks.mass_h1(c, NC, v, z, want_den && ks.tuned_mass); }
      if (rc) { return rc; }
      if (c.profile_mass) { int rt = timer_end(c, 4); if (rt) { return rt; } c.mass_launches++; }
      if (want_den && (ks.tuned_mass || bapply)) { den_blocks = c.dt_nblocks; den_src_off = bapply ? 0 : c.den_off; }
      return halo_sum(c, z, NC);
   };
   // per-block partials -> (sum over ranks) -> what the finish kernels read.
   // Single rank: the finish kernel reduces the partials itself (fixed order).
   // Multi rank: reduce to NC sums, NCCL all-reduce in-stream, finish reads the NC sums.
   // Multi rank, peer memory (device/p2p.cuh): the finish kernel itself publishes its NC sums to every rank and adds
   // the ranks' values in rank order (pd != nullptr).  Multi rank, NCCL: reduce to NC sums, all-reduce in-stream,
   // the finish kernel reads the NC sums.
   const p2p::Dev *pd = nullptr; unsigned long long pseq = 0;
   auto reduced = [&](int nblocks, double *tmp, const double *&src, int &nsrc, size_t off = 0) -> int
   {
      pd = nullptr; pseq = 0;
      const double *part = c.d_part + off;
      if (c.nranks <= 1) { src = part; nsrc = nblocks; return LAGB_OK; }
      if (c.p2p_on && c.tune[11] == 0) { src = part; nsrc = nblocks; pd = c.d_p2p_dev; pseq = ++c.p2p_scal_seq; return LAGB_OK; }
      src = tmp; nsrc = 1;
      LAGB_LAUNCH_K(c, pcg::reduce_final<NC>, 1, pcg::FB, 0, nblocks, part, tmp);
      return allreduce_sum(c, tmp, NC);
   };

   int rc, den_blocks = 0, nsrc = 0;
   const double *src = nullptr;
   if (iterative_mode)
   {
      if (zero_z && !zalt) { LAGB_CUDA(cudaMemsetAsync(z, 0, sizeof(double)*NC*n, c.stream)); }
      rc = apply(x, false, den_blocks); if (rc) { return rc; }
   }
   else { LAGB_CUDA(cudaMemsetAsync(x, 0, sizeof(double)*NC*n, c.stream)); }
   // bit 0: r = b - A x ; bit 1: clear z (not needed when the alternating buffers are clean and no A x was formed)
   LAGB_LAUNCH_K(c, pcg::init_residual<NC>, g, pcg::RB, 0, n, cs, b, z, P, own, r, d, c.d_part,
                 (iterative_mode ? 1 : 0) | ((iterative_mode || !zalt) ? 2 : 0));
   rc = reduced(g, c.d_tmp, src, nsrc); if (rc) { return rc; }
   LAGB_LAUNCH_K(c, pcg::finish_init<NC>, fin_grid(nsrc), pcg::FB, 0, c.d_state, src, nsrc, rel_tol, 0.0, pd, pseq, c.d_fin, c.d_grp_ctr);

   // The host only needs to know when every component has stopped; iterations are
   // enqueued ahead (kernels skip finished components) and the flag is polled at
   // a cadence derived from the previous solve's iteration count.
   int next_check = std::max(1, c.predicted_iters - 1);
   bool finished = false;
   int it = 0;
   auto poll = [&]() -> int
   {
      LAGB_CUDA(cudaMemcpyAsync(c.h_state, c.d_state, sizeof(pcg::State), cudaMemcpyDeviceToHost, c.stream));
      LAGB_CUDA(cudaStreamSynchronize(c.stream));
      finished = c.h_state->all_done != 0;
      return LAGB_OK;
   };
   if (c.predicted_iters == 0) { rc = poll(); if (rc) { return rc; } }
   while (!finished && it < max_iter)
   {
      it++;
      if (zalt) { rc = z_acquire(zi); if (rc) { return rc; } z = zb[zi]; }
      rc = apply(d, true, den_blocks); if (rc) { return rc; }
      if (den_blocks == 0)
      {
         LAGB_LAUNCH_K(c, pcg::dot_partial<NC>, g, pcg::RB, 0, n, cs, (const double*)d, (const double*)z, own, c.d_part);
         den_blocks = g; den_src_off = 0;
      }
      rc = reduced(den_blocks, c.d_tmp, src, nsrc, den_src_off); if (rc) { return rc; }
      LAGB_LAUNCH_K(c, pcg::finish_den<NC>, fin_grid(nsrc), pcg::FB, 0, c.d_state, src, nsrc, it, pd, pseq, c.d_fin, c.d_grp_ctr);
      if (c.tune[8] == 1)
      {
         // first version: x and r in one kernel, d (and the zero fill of z) in another
         LAGB_LAUNCH_K(c, pcg::update_xr<NC>, g, pcg::RB, 0, n, cs, (const pcg::State*)c.d_state, x, r, (const double*)d, (const double*)z, P, own, c.d_part);
         rc = reduced(g, c.d_tmp + 4, src, nsrc); if (rc) { return rc; }
         LAGB_LAUNCH_K(c, pcg::finish_beta<NC>, fin_grid(nsrc), pcg::FB, 0, c.d_state, src, nsrc, it, max_iter, pd, pseq, c.d_fin, c.d_grp_ctr);
         LAGB_LAUNCH_K(c, pcg::update_d<NC>, g, pcg::RB, 0, n, cs, (const pcg::State*)c.d_state, d, (const double*)r, P, z);
      }
      else
      {
         // same arithmetic, one vector pass less: x is updated where d is read anyway
         LAGB_LAUNCH_K(c, pcg::update_r<NC>, g_r, pcg::RB, 0, n, cs, (const pcg::State*)c.d_state, r, (const double*)z, P, own, c.d_part);
         if (zalt) { rc = z_release(zi); if (rc) { return rc; } zi ^= 1; }
         rc = reduced(g_r, c.d_tmp + 4, src, nsrc); if (rc) { return rc; }
         LAGB_LAUNCH_K(c, pcg::finish_beta<NC>, fin_grid(nsrc), pcg::FB, 0, c.d_state, src, nsrc, it, max_iter, pd, pseq, c.d_fin, c.d_grp_ctr);
         auto kz = pcg::update_dx<NC,true>; auto kn = pcg::update_dx<NC,false>;
         LAGB_LAUNCH_K(c, (zero_z && !zalt) ? kz : kn, g_dx, pcg::RB, 0, n, cs, (const pcg::State*)c.d_state, x, d, (const double*)r, P, z);
      }
      if (it >= next_check) { rc = poll(); if (rc) { return rc; } }
   }
   if (!finished) { rc = poll(); if (rc) { return rc; } }
   int mx = 0;
   for (int k = 0; k < NC; k++) { h_iters[k] = c.h_state->iters[k]; mx = std::max(mx, h_iters[k]); }
   c.predicted_iters = mx;
   return LAGB_OK;
}

// ---------------------------------------------------------------------------
// The same solver on the brick schedule (device/mass3d_brick.cuh): the direction update
// d <- M^-1 r + beta d is formed inside the operator's gather (two buffers, swapped every
// iteration), the operator writes with plain stores (no zero fill of A d), and the vector
// work per iteration is the single pass update_xr.  Iterates are those of pcg_run_nc.
// ---------------------------------------------------------------------------
template<int NC>
static int pcg_run_brick(Ctx &c, int comp0, const double *b, double *x, double rel_tol, int max_iter,
                         bool iterative_mode, int *h_iters)
{
   const int64_t n = c.ndofs, cs = n;
   double *r = c.d_r, *z = c.d_z;
   double *dcur = c.d_d, *dnxt = c.d_d2;
   c.z_clean = false;     // both buffers are scratch here (pcg_run_nc restores its zero invariant on its next solve)
   if (c.aux_stream) { LAGB_CUDA(cudaStreamSynchronize(c.aux_stream)); c.z_pending[0] = c.z_pending[1] = false; }
   pcg::Prec P; P.dinv = c.d_dinv; P.ess = c.d_essmask; P.comp0 = comp0;
   const unsigned char *own = c.d_own;
   const int g = vec_grid(n);
   if (g*NC > c.part_cap) { set_error("pcg: partial buffer too small"); return LAGB_ERR_STATE; }
   KernelSet &ks = c.ks;
   auto reduced = [&](int nblocks, double *tmp, const double *&src, int &nsrc) -> int
   {
      if (c.nranks <= 1) { src = c.d_part; nsrc = nblocks; return LAGB_OK; }
      src = tmp; nsrc = 1;
      if (c.p2p_on && c.tune[11] == 0)
      {
         // one launch: last-stage reduction + publication to every rank + fixed-rank-order sum
         p2p::p2p_allreduce<NC><<<1, pcg::FB, 0, c.stream>>>(c.p2p_dev, ++c.p2p_scal_seq, nblocks, c.d_part, tmp);
         LAGB_LAUNCH_CHECK();
         return LAGB_OK;
      }
      pcg::reduce_final<NC><<<1, pcg::FB, 0, c.stream>>>(nblocks, c.d_part, tmp);
      LAGB_LAUNCH_CHECK();
      return allreduce_sum(c, tmp, NC);
   };
   auto apply = [&](const MassBrickIn &in, bool want_den) -> int
   {
      if (c.profile_mass) { int rt = timer_begin(c, 4); if (rt) { return rt; } }
      int rc = ks.mass_brick(c, NC, in, z, want_den); if (rc) { return rc; }
      if (c.profile_mass) { int rt = timer_end(c, 4); if (rt) { return rt; } c.mass_launches++; }
      return halo_sum(c, z, NC);
   };
   int rc, nsrc = 0;
   const double *src = nullptr;
   if (iterative_mode) { MassBrickIn in; in.x = x; rc = apply(in, false); if (rc) { return rc; } }
   else { LAGB_CUDA(cudaMemsetAsync(x, 0, sizeof(double)*NC*n, c.stream)); }
   pcg::init_residual<NC><<<g, pcg::RB, 0, c.stream>>>(n, cs, b, z, P, own, r, dcur, c.d_part, (iterative_mode ? 1 : 0) | 2);
   LAGB_LAUNCH_CHECK();
   rc = reduced(g, c.d_tmp, src, nsrc); if (rc) { return rc; }
   pcg::finish_init<NC><<<1, pcg::FB, 0, c.stream>>>(c.d_state, src, nsrc, rel_tol, 0.0, nullptr, 0ull, c.d_fin, c.d_grp_ctr);   // beta = 0: first direction = M^-1 r
   LAGB_LAUNCH_CHECK();

   int next_check = std::max(1, c.predicted_iters - 1);
   bool finished = false;
   int it = 0;
   auto poll = [&]() -> int
   {
      LAGB_CUDA(cudaMemcpyAsync(c.h_state, c.d_state, sizeof(pcg::State), cudaMemcpyDeviceToHost, c.stream));
      LAGB_CUDA(cudaStreamSynchronize(c.stream));
      finished = c.h_state->all_done != 0;
      return LAGB_OK;
   };
   if (c.predicted_iters == 0) { rc = poll(); if (rc) { return rc; } }
   while (!finished && it < max_iter)
   {
      it++;
      MassBrickIn in; in.r = r; in.dold = dcur; in.dnew = dnxt; in.comp0 = comp0;
      rc = apply(in, true); if (rc) { return rc; }
      rc = reduced(c.dt_nblocks, c.d_tmp, src, nsrc); if (rc) { return rc; }
      pcg::finish_den<NC><<<1, pcg::FB, 0, c.stream>>>(c.d_state, src, nsrc, it, nullptr, 0ull, c.d_fin, c.d_grp_ctr);
      LAGB_LAUNCH_CHECK();
      pcg::update_xr<NC><<<g, pcg::RB, 0, c.stream>>>(n, cs, c.d_state, x, r, dnxt, z, P, own, c.d_part);
      LAGB_LAUNCH_CHECK();
      rc = reduced(g, c.d_tmp + 4, src, nsrc); if (rc) { return rc; }
      pcg::finish_beta<NC><<<1, pcg::FB, 0, c.stream>>>(c.d_state, src, nsrc, it, max_iter, nullptr, 0ull, c.d_fin, c.d_grp_ctr);
      LAGB_LAUNCH_CHECK();
      std::swap(dcur, dnxt);
      if (it >= next_check) { rc = poll(); if (rc) { return rc; } }
   }
   if (!finished) { rc = poll(); if (rc) { return rc; } }
   int mx = 0;
   for (int k = 0; k < NC; k++) { h_iters[k] = c.h_state->iters[k]; mx = std::max(mx, h_iters[k]); }
   c.predicted_iters = mx;
   return LAGB_OK;
}

// tune[6]: 0 = default path, 1 = legacy atomic scatter, 2 = brick v1, 3 = brick v2
static bool use_brick(const Ctx &c) { return c.variant == 0 && c.ks.mass_brick != nullptr && c.tune[6] >= 2; }

static int pcg_run(Ctx &c, bool l2, int nc, int comp0, const double *b, double *x, double rel_tol,
                   int max_iter, bool iterative_mode, int *h_iters)
{
   if (!l2 && use_brick(c) && c.tune[6] <= 3 && c.tune[9] == 0)   // fused direction update (multi-launch brick kernels)
   {
      if (nc == 1) { return pcg_run_brick<1>(c, comp0, b, x, rel_tol, max_iter, iterative_mode, h_iters); }
      if (nc == 3) { return pcg_run_brick<3>(c, comp0, b, x, rel_tol, max_iter, iterative_mode, h_iters); }
   }
   switch (nc)
   {
      case 1: return pcg_run_nc<1>(c, l2, comp0, b, x, rel_tol, max_iter, iterative_mode, h_iters);
      case 2: return pcg_run_nc<2>(c, l2, comp0, b, x, rel_tol, max_iter, iterative_mode, h_iters);
      case 3: return pcg_run_nc<3>(c, l2, comp0, b, x, rel_tol, max_iter, iterative_mode, h_iters);
   }
   set_error("pcg: bad component count"); return LAGB_ERR_INVALID;
}

} // namespace lagb

using namespace lagb;

// every entry point that takes a context runs on the context's device (the caller may have changed
// the current device between calls; one process can drive several contexts)
static inline int enter(Ctx &c)
{
   int cur = -1;
   if (cudaGetDevice(&cur) != cudaSuccess || cur != c.device) { LAGB_CUDA(cudaSetDevice(c.device)); }
   return LAGB_OK;
}
#define LAGB_ENTER(h) do { if (!(h)) { set_error("null context"); return LAGB_ERR_INVALID; } int e__ = enter((h)->c); if (e__) { return e__; } } while (0)

extern "C" {

const char *lagb_last_error(void) { return g_err.c_str(); }
int64_t lagb_kernel_launch_count(void) { return g_launch_count; }

struct lagb_ctx { Ctx c; };

int lagb_ctx_create(lagb_ctx **out, const lagb_ctx_desc *d, void *stream)
{
   if (!out || !d) { set_error("ctx_create: null argument"); return LAGB_ERR_INVALID; }
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
   { set_error("no CUDA device: laghos_b200 has no CPU fallback"); return LAGB_ERR_CUDA; }
   LAGB_CUDA(cudaSetDevice(d->device));
   lagb_ctx *h = new lagb_ctx();
   Ctx &c = h->c;
   c.dim = d->dim; c.NE = d->NE; c.D1D = d->D1D; c.L1D = d->L1D; c.Q1D = d->Q1D;
   if (c.L1D != c.D1D - 1) { set_error("L1D!=D1D-1"); delete h; return LAGB_ERR_INVALID; } // reference laghos_assembly.cpp:533
   c.ND = ipow(c.D1D, c.dim); c.NL = ipow(c.L1D, c.dim); c.NQ = ipow(c.Q1D, c.dim);
   c.ndofs = d->ndofs_h1; c.ndofs_l2 = (int64_t)c.NE*c.NL;
   c.use_visc = d->use_visc; c.use_vort = d->use_vort; c.variant = d->kernel_variant; c.device = d->device;
   c.stream = (cudaStream_t)stream;
   {
      int sms = 0;
      if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d->device) != cudaSuccess)
      { set_error("cudaDeviceGetAttribute(multiProcessorCount) failed"); delete h; return LAGB_ERR_CUDA; }
      c.num_sms = sms;
   }
   c.ks_generic = make_generic_kernels(c.dim, c.D1D, c.Q1D);
   if (!c.ks_generic.mass_h1)
   {
      char msg[64]; snprintf(msg, sizeof msg, "Unknown kernel 0x%x", (c.dim << 8) | (c.D1D << 4) | c.Q1D);
      set_error(msg); delete h; return LAGB_ERR_INVALID;
   }
   c.ks = c.ks_generic;
   add_tuned_kernels(c.ks, c.dim, c.D1D, c.Q1D);
   // tables blob: B | G | BL in DevTables<D1D,Q1D> order
   const int nb = c.Q1D*c.D1D, nbl = c.Q1D*std::max(1, c.L1D);
   c.tab_blob.resize(sizeof(double)*(2*nb + nbl));
   memcpy(c.tab_blob.data(), d->h_B, sizeof(double)*nb);
   memcpy(c.tab_blob.data() + sizeof(double)*nb, d->h_G, sizeof(double)*nb);
   memcpy(c.tab_blob.data() + sizeof(double)*2*nb, d->h_BL, sizeof(double)*c.Q1D*c.L1D);
   int rc = 0;
   const size_t NEQ = (size_t)c.NE*c.NQ, D2 = (size_t)c.dim*c.dim;
   rc |= dev_upload(&c.d_map, d->h_h1_map, (size_t)c.NE*c.ND);
   c.h_map.assign(d->h_h1_map, d->h_h1_map + (size_t)c.NE*c.ND);
   for (int k = 0; k < 3; k++) { c.elem_grid[k] = d->elem_grid[k]; }
   for (int k = 0; k < c.dim; k++) { c.ness[k] = d->ness[k]; rc |= dev_upload(&c.d_ess[k], d->h_ess[k], (size_t)d->ness[k]); }
   rc |= dev_upload(&c.d_qweights, d->h_qweights, (size_t)c.NQ);
   {
      std::vector<double> iw(c.NQ);
      for (int q = 0; q < c.NQ; q++) { iw[q] = 1.0/d->h_qweights[q]; }
      rc |= dev_upload(&c.d_inv_qweights, iw.data(), (size_t)c.NQ);
   }
   rc |= dev_upload(&c.d_gamma, d->h_gamma, (size_t)c.NE);
   rc |= dev_alloc(&c.d_sJit, NEQ*D2); rc |= dev_alloc(&c.d_rho0DetJ0w, NEQ);
   rc |= dev_alloc(&c.d_Jac0inv, NEQ*D2); rc |= dev_alloc(&c.d_massD, NEQ);
   rc |= dev_alloc(&c.d_diag, (size_t)c.ndofs); rc |= dev_alloc(&c.d_dinv, (size_t)c.ndofs);
   {
      // bit c of essmask[i]: scalar dof i is essential for velocity component c
      std::vector<unsigned char> em((size_t)c.ndofs, 0);
      for (int k = 0; k < c.dim; k++) { for (int j = 0; j < d->ness[k]; j++) { em[d->h_ess[k][j]] |= (unsigned char)(1u << k); } }
      rc |= dev_upload(&c.d_essmask, em.data(), em.size());
   }
   rc |= dev_alloc(&c.d_r, (size_t)c.ndofs*c.dim); rc |= dev_alloc(&c.d_d, (size_t)c.ndofs*c.dim);
   rc |= dev_alloc(&c.d_z, (size_t)c.ndofs*c.dim); rc |= dev_alloc(&c.d_d2, (size_t)c.ndofs*c.dim);
   rc |= dev_alloc(&c.d_lr, (size_t)c.ndofs_l2); rc |= dev_alloc(&c.d_ld, (size_t)c.ndofs_l2);
   rc |= dev_alloc(&c.d_lz, (size_t)c.ndofs_l2);
   c.part_cap = std::max(c.NE, 148*16)*4 + 64;
   rc |= dev_alloc(&c.d_part, (size_t)c.part_cap);
   c.grp_cap = 64;
   rc |= dev_alloc(&c.d_grp_ctr, (size_t)c.grp_cap);
   rc |= dev_alloc(&c.d_fin, (size_t)pcg::FIN_CTAS*pcg::MAXC);
   if (!rc) { rc |= (cudaMemset(c.d_grp_ctr, 0, sizeof(unsigned int)*(size_t)c.grp_cap) != cudaSuccess); }
   rc |= dev_alloc(&c.d_tmp, 16); rc |= dev_alloc(&c.d_dt, 1); rc |= dev_alloc(&c.d_elem_vol, (size_t)c.NE);
   rc |= dev_alloc(&c.d_state, 1);
   if (rc) { lagb_ctx_destroy(h); return LAGB_ERR_CUDA; }
   // late failures release what was allocated so far (lagb_ctx_destroy accepts a partly built context)
   auto finish = [&]() -> int
   {
      LAGB_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
      LAGB_CUDA(cudaStreamCreateWithFlags(&c.aux_stream, cudaStreamNonBlocking));
      LAGB_CUDA(cudaEventCreateWithFlags(&c.ev_zuse, cudaEventDisableTiming));
      LAGB_CUDA(cudaEventCreateWithFlags(&c.ev_zclr[0], cudaEventDisableTiming));
      LAGB_CUDA(cudaEventCreateWithFlags(&c.ev_zclr[1], cudaEventDisableTiming));
      LAGB_CUDA(cudaMemset(c.d_z, 0, sizeof(double)*(size_t)c.ndofs*c.dim));
      LAGB_CUDA(cudaMemset(c.d_d2, 0, sizeof(double)*(size_t)c.ndofs*c.dim));
      LAGB_CUDA(cudaEventCreateWithFlags(&c.ev_compute, cudaEventDisableTiming));
      LAGB_CUDA(cudaEventCreateWithFlags(&c.ev_copy, cudaEventDisableTiming));
      LAGB_CUDA(cudaMallocHost((void**)&c.h_state, sizeof(pcg::State)));
      LAGB_CUDA(cudaMallocHost((void**)&c.h_scal, 16*sizeof(double)));
      LAGB_CUDA(cudaMemset(c.d_sJit, 0, NEQ*D2*sizeof(double)));
      return LAGB_OK;
   };
   rc = finish();
   if (rc) { lagb_ctx_destroy(h); return rc; }
   // measurement aid: LAGB_TUNE="key=value,key=value" presets lagb_tune_set for contexts created by drivers that
   // do not expose the knobs (bench.py A/B runs)
   if (const char *tv = getenv("LAGB_TUNE"))
   {
      int k = 0, v = 0, used = 0;
      while (sscanf(tv, " %d=%d%n", &k, &v, &used) == 2)
      {
         if (k >= 0 && k < 16) { c.tune[k] = v; }
         tv += used;
         if (*tv == ',') { tv++; }
      }
   }
   *out = h;
   return LAGB_OK;
}

void lagb_ctx_destroy(lagb_ctx *h)
{
   if (!h) { return; }
   Ctx &c = h->c;
   cudaStreamSynchronize(c.stream);
   void *ptrs[] = {c.d_map, c.d_ess[0], c.d_ess[1], c.d_ess[2], c.d_qweights, c.d_inv_qweights, c.d_gamma, c.d_sJit, c.d_rho0DetJ0w,
                   c.d_Jac0inv, c.d_massD, c.d_diag, c.d_dinv, c.d_essmask, c.d_r, c.d_d, c.d_z, c.d_lr, c.d_ld, c.d_lz,
                   c.d_part, c.d_tmp, c.d_dt, c.d_elem_vol, c.d_state, c.d_own, c.d_d2, c.d_l2inv, c.d_BL, c.d_grp_ctr, c.d_fin};
   for (void *p : ptrs) { if (p) { cudaFree(p); } }
   for (auto &kv : c.plans)
   {
      DevPlan &dp = kv.second;
      void *pp[] = {dp.belem, dp.bnuniq, dp.btab, dp.buid, dp.lidx, dp.uoff, dp.upos, dp.ucon, dp.bmeta, dp.deps, dp.flags, dp.work_ctr};
      for (void *p : pp) { if (p) { cudaFree(p); } }
   }
   for (auto &nb : c.nbrs) { cudaFree(nb.d_idx); cudaFree(nb.d_send); cudaFree(nb.d_recv); }
   void *hp[] = {c.d_pack_idx, c.d_pack_nb, c.d_nbr_off, c.d_nbr_n, c.d_u_dof, c.d_u_ptr, c.d_u_src, c.d_send_all, c.d_recv_all};
   for (void *p : hp) { if (p) { cudaFree(p); } }
   if (c.copy_stream) { cudaStreamSynchronize(c.copy_stream); cudaStreamDestroy(c.copy_stream); }
   if (c.aux_stream) { cudaStreamSynchronize(c.aux_stream); cudaStreamDestroy(c.aux_stream); }
   { cudaEvent_t ev[] = {c.ev_zuse, c.ev_zclr[0], c.ev_zclr[1]}; for (cudaEvent_t e : ev) { if (e) { cudaEventDestroy(e); } } }
   if (c.ev_compute) { cudaEventDestroy(c.ev_compute); }
   if (c.ev_copy) { cudaEventDestroy(c.ev_copy); }
   if (c.h_state) { cudaFreeHost(c.h_state); }
   if (c.h_scal) { cudaFreeHost(c.h_scal); }
   for (int w = 0; w < Timer::NT; w++) { for (auto &p : c.timer.pending[w]) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); } }
   for (auto e : c.timer.pool) { cudaEventDestroy(e); }
   for (void *p : c.p2p_opened) { cudaIpcCloseMemHandle(p); }
   { void *pp[] = {c.p2p_base, c.d_nbr_roff, c.d_nbr_rank, c.d_pack_done, c.d_p2p_dev}; for (void *p : pp) { if (p) { cudaFree(p); } } }
   if (c.nccl_comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(c.nccl_comm); }
   delete h;
}

int lagb_ctx_sync(lagb_ctx *h)
{
   LAGB_ENTER(h);
   LAGB_CUDA(cudaStreamSynchronize(h->c.stream));
   if (h->c.copy_stream) { LAGB_CUDA(cudaStreamSynchronize(h->c.copy_stream)); }
   if (h->c.aux_stream) { LAGB_CUDA(cudaStreamSynchronize(h->c.aux_stream)); }
   return LAGB_OK;
}

int lagb_setup_qdata0(lagb_ctx *h, const double *d_x0, const double *d_rho0_gf, const double *d_rho0_q,
                      int64_t ne_global, double *h0_out)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   KernelSet &ks = (c.variant == 1) ? c.ks_generic : c.ks;
   int rc = ks.rho0detj0(c, d_x0, d_rho0_gf, d_rho0_q, c.d_elem_vol); if (rc) { return rc; }
   // volume = vol * one (reference laghos_solver.cpp:1196-1260): ONE serial sum over all NE*NQ values of
   // w detJ0 in the reference's order (element outer, point inner), on the host, once.  The sum's round-off
   // (2e-11 relative in h0 at 32^3 elements, 1e-10 at 64^3) enters every viscosity coefficient through h0,
   // so a better-conditioned summation would move |e| away from the reference's serial CPU path by that much.
   double vol = 0.0;
   {
      const size_t NEQ = (size_t)c.NE*c.NQ;
      rc = c.ks_generic.detj_w(c, d_x0, c.d_sJit); if (rc) { return rc; }      // stressJinvT is not in use yet: scratch
      std::vector<double> wd(NEQ);
      LAGB_CUDA(cudaMemcpyAsync(wd.data(), c.d_sJit, sizeof(double)*NEQ, cudaMemcpyDeviceToHost, c.stream));
      LAGB_CUDA(cudaMemsetAsync(c.d_sJit, 0, sizeof(double)*NEQ, c.stream));
      LAGB_CUDA(cudaStreamSynchronize(c.stream));
      for (size_t i = 0; i < NEQ; i++) { vol += wd[i]; }
   }
   double ne = (double)c.NE;
   if (c.nranks > 1)
   {
      c.h_scal[0] = vol; c.h_scal[1] = ne;
      LAGB_CUDA(cudaMemcpyAsync(c.d_tmp, c.h_scal, 2*sizeof(double), cudaMemcpyHostToDevice, c.stream));
      rc = allreduce_sum(c, c.d_tmp, 2); if (rc) { return rc; }
      LAGB_CUDA(cudaMemcpyAsync(c.h_scal, c.d_tmp, 2*sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      LAGB_CUDA(cudaStreamSynchronize(c.stream));
      vol = c.h_scal[0]; ne = c.h_scal[1];
   }
   if (ne_global > 0) { ne = (double)ne_global; }
   c.ne_global = (int64_t)ne;
   // reference laghos_solver.cpp:253-262 (SQUARE / CUBE)
   c.h0 = (c.dim == 2) ? sqrt(vol/ne) : pow(vol/ne, 1./3.);
   c.h0 /= (double)(c.D1D - 1);
   // Jacobi diagonal (MFEM AssembleDiagonalPA) and its masked inverse per component
   LAGB_CUDA(cudaMemsetAsync(c.d_diag, 0, sizeof(double)*c.ndofs, c.stream));
   rc = ks.mass_diag(c, c.d_diag); if (rc) { return rc; }
   rc = halo_sum(c, c.d_diag, 1); if (rc) { return rc; }
   build_dinv<<<vec_grid(c.ndofs), pcg::RB, 0, c.stream>>>(c.ndofs, c.d_diag, c.d_dinv);
   LAGB_LAUNCH_CHECK();
   LAGB_CUDA(cudaStreamSynchronize(c.stream));
   c.setup_done = true;
   if (h0_out) { *h0_out = c.h0; }
   return LAGB_OK;
}

int lagb_vmass_mult(lagb_ctx *h, int comp, const double *d_x, double *d_y)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (comp >= c.dim) { set_error("vmass_mult: bad component"); return LAGB_ERR_INVALID; }
   KernelSet &ks = (c.variant == 1) ? c.ks_generic : c.ks;
   int rc;
   if (use_brick(c)) { MassBrickIn in; in.x = d_x; rc = ks.mass_brick(c, 1, in, d_y, false); if (rc) { return rc; } }
   else
   {
      LAGB_CUDA(cudaMemsetAsync(d_y, 0, sizeof(double)*c.ndofs, c.stream));
      rc = ks.mass_h1(c, 1, d_x, d_y, false); if (rc) { return rc; }
   }
   rc = halo_sum(c, d_y, 1); if (rc) { return rc; }
   if (comp >= 0 && c.ness[comp] > 0)
   {
      pcg::vec_zero_idx<<<std::max(1, std::min(1024, (c.ness[comp] + 255)/256)), 256, 0, c.stream>>>(d_y, c.d_ess[comp], c.ness[comp]);
      LAGB_LAUNCH_CHECK();
   }
   return LAGB_OK;
}

int lagb_vmass_mult_all(lagb_ctx *h, const double *d_x, double *d_y)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   KernelSet &ks = (c.variant == 1) ? c.ks_generic : c.ks;
   int rc;
   if (use_brick(c) && c.dim == 3) { MassBrickIn in; in.x = d_x; rc = ks.mass_brick(c, 3, in, d_y, false); if (rc) { return rc; } }
   else
   {
      LAGB_CUDA(cudaMemsetAsync(d_y, 0, sizeof(double)*c.ndofs*c.dim, c.stream));
      rc = ks.mass_h1(c, c.dim, d_x, d_y, false); if (rc) { return rc; }
   }
   return halo_sum(c, d_y, c.dim);
}

int lagb_host_batch_plan_check(const int32_t *h_map, int NE, int ND, int64_t ndofs, const int32_t grid[3], int NB, int64_t stats[8])
{
   BatchPlan bp; std::string err;
   const int g[3] = {grid ? grid[0] : 0, grid ? grid[1] : 0, grid ? grid[2] : 0}, shape[3] = {0, 0, 0};
   if (bp.build(h_map, NE, ND, ndofs, g, NB, shape, err)) { set_error(err); return LAGB_ERR_INVALID; }
   if (bp.self_check(h_map, NE, ndofs, err)) { set_error("batch plan: " + err); return LAGB_ERR_STATE; }
   if (stats)
   {
      stats[0] = bp.nbatch; stats[1] = bp.ncolors; stats[2] = bp.ntab; stats[3] = bp.umax; stats[4] = bp.UP;
      stats[5] = bp.n_first; stats[6] = bp.brick[0] + 100*bp.brick[1] + 10000*bp.brick[2];
      int64_t tot = 0; for (int u : bp.nuniq) { tot += u; }
      stats[7] = tot;
   }
   return LAGB_OK;
}

int lagb_tune_set(lagb_ctx *h, int key, int value)
{
   if (key < 0 || key >= 16) { set_error("tune_set: bad key"); return LAGB_ERR_INVALID; }
   h->c.tune[key] = value;
   return LAGB_OK;
}

int lagb_vmass_diag(lagb_ctx *h, double *d_diag)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (!c.setup_done) { set_error("vmass_diag: call lagb_setup_qdata0 first"); return LAGB_ERR_STATE; }
   LAGB_CUDA(cudaMemcpyAsync(d_diag, c.d_diag, sizeof(double)*c.ndofs, cudaMemcpyDeviceToDevice, c.stream));
   return LAGB_OK;
}

int lagb_emass_mult(lagb_ctx *h, const double *d_x, double *d_y)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   KernelSet &ks = (c.variant == 1) ? c.ks_generic : c.ks;
   return ks.mass_l2(c, d_x, d_y);
}

int lagb_force_mult(lagb_ctx *h, const double *d_e, double *d_v)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   KernelSet &ks = (c.variant == 1) ? c.ks_generic : c.ks;
   int rc = timer_begin(c, 2); if (rc) { return rc; }
   LAGB_CUDA(cudaMemsetAsync(d_v, 0, sizeof(double)*c.ndofs*c.dim, c.stream));
   rc = ks.force_mult(c, d_e, d_v); if (rc) { return rc; }
   rc = halo_sum(c, d_v, c.dim); if (rc) { return rc; }
   return timer_end(c, 2);
}

int lagb_force_mult_transpose(lagb_ctx *h, const double *d_v, double *d_e)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   KernelSet &ks = (c.variant == 1) ? c.ks_generic : c.ks;
   int rc = timer_begin(c, 2); if (rc) { return rc; }
   rc = ks.force_mult_t(c, d_v, d_e); if (rc) { return rc; }
   return timer_end(c, 2);
}

int lagb_dt_est_set(lagb_ctx *h, double v)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   pcg::vec_fill<<<1, 32, 0, c.stream>>>(c.d_dt, v, 1);
   LAGB_LAUNCH_CHECK();
   return LAGB_OK;
}

int lagb_qupdate_async(lagb_ctx *h, const double *d_S, double cfl)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (!c.setup_done) { set_error("qupdate: call lagb_setup_qdata0 first"); return LAGB_ERR_STATE; }
   KernelSet &ks = (c.variant == 1) ? c.ks_generic : c.ks;
   QPointParams prm;
   prm.h0 = c.h0; prm.h1order = (double)(c.D1D - 1); prm.inv_h1order = 1.0/prm.h1order; prm.cfl = cfl; prm.dt_in = std::numeric_limits<double>::infinity();
   prm.use_viscosity = c.use_visc; prm.use_vorticity = c.use_vort;
   int rc = timer_begin(c, 3); if (rc) { return rc; }
   rc = ks.qupdate(c, d_S, prm); if (rc) { return rc; }
   rc = timer_end(c, 3); if (rc) { return rc; }
   c.quad_tstep += c.NE;
   return LAGB_OK;
}

int lagb_dt_est_read(lagb_ctx *h, double *out)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   int rc = allreduce_min(c, c.d_dt, 1); if (rc) { return rc; }
   LAGB_CUDA(cudaMemcpyAsync(c.h_scal, c.d_dt, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
   LAGB_CUDA(cudaStreamSynchronize(c.stream));
   *out = c.h_scal[0];
   return LAGB_OK;
}

int lagb_qupdate(lagb_ctx *h, const double *d_S, double cfl, double dt_est_in, double *out)
{
   LAGB_ENTER(h);
   int rc = lagb_dt_est_set(h, dt_est_in); if (rc) { return rc; }
   rc = lagb_qupdate_async(h, d_S, cfl); if (rc) { return rc; }
   return lagb_dt_est_read(h, out);
}

int lagb_pcg_vmass(lagb_ctx *h, int comp, const double *d_b, double *d_x, double rel_tol, int max_iter, int *iters)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (!c.setup_done) { set_error("pcg_vmass: call lagb_setup_qdata0 first"); return LAGB_ERR_STATE; }
   if (comp < 0 || comp >= c.dim) { set_error("pcg_vmass: bad component"); return LAGB_ERR_INVALID; }
   int rc = timer_begin(c, 0); if (rc) { return rc; }
   int it = 0;
   rc = pcg_run(c, false, 1, comp, d_b, d_x, rel_tol, max_iter, true, &it); if (rc) { return rc; }
   rc = timer_end(c, 0); if (rc) { return rc; }
   c.H1iter += it;
   if (iters) { *iters = it; }
   return LAGB_OK;
}

int lagb_pcg_vmass_all(lagb_ctx *h, const double *d_rhs, double *d_dv, double rel_tol, int max_iter, int *iters)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (!c.setup_done) { set_error("pcg_vmass_all: call lagb_setup_qdata0 first"); return LAGB_ERR_STATE; }
   int rc = timer_begin(c, 0); if (rc) { return rc; }
   int it[3] = {0, 0, 0};
   rc = pcg_run(c, false, c.dim, 0, d_rhs, d_dv, rel_tol, max_iter, true, it); if (rc) { return rc; }
   rc = timer_end(c, 0); if (rc) { return rc; }
   for (int k = 0; k < c.dim; k++) { c.H1iter += it[k]; if (iters) { iters[k] = it[k]; } }
   return LAGB_OK;
}

int lagb_pcg_vmass_all_x0(lagb_ctx *h, const double *d_rhs, double *d_dv, double rel_tol, int max_iter, int *iters)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (!c.setup_done) { set_error("pcg_vmass_all_x0: call lagb_setup_qdata0 first"); return LAGB_ERR_STATE; }
   int rc = timer_begin(c, 0); if (rc) { return rc; }
   int it[3] = {0, 0, 0};
   rc = pcg_run(c, false, c.dim, 0, d_rhs, d_dv, rel_tol, max_iter, false, it); if (rc) { return rc; }
   rc = timer_end(c, 0); if (rc) { return rc; }
   for (int k = 0; k < c.dim; k++) { c.H1iter += it[k]; if (iters) { iters[k] = it[k]; } }
   return LAGB_OK;
}

int lagb_cg_emass(lagb_ctx *h, const double *d_b, double *d_x, double rel_tol, int max_iter, int *iters)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (!c.setup_done) { set_error("cg_emass: call lagb_setup_qdata0 first"); return LAGB_ERR_STATE; }
   int rc = timer_begin(c, 1); if (rc) { return rc; }
   int it = 0;
   // element inverses unless the CG is asked for (lagb_tune_set key 10 / LAGB_L2_SOLVER=cg):
   // reported as 0 iterations, which the reference counts as one (laghos_solver.cpp:485-486)
   bool direct = false;
   rc = l2_direct_solve(c, d_b, d_x, &direct); if (rc) { return rc; }
   if (!direct)
   {
      const int pred = c.predicted_iters;
      c.predicted_iters = 0;
      rc = pcg_run(c, true, 1, 0, d_b, d_x, rel_tol, max_iter, false, &it);
      c.predicted_iters = pred;
      if (rc) { return rc; }
   }
   rc = timer_end(c, 1); if (rc) { return rc; }
   c.L2iter += (it == 0) ? 1 : it;   // reference laghos_solver.cpp:485-486
   if (iters) { *iters = it; }
   return LAGB_OK;
}

// sum over ranks of dot(a, b) (n entries); scratch: d_part, d_tmp + 8
static int global_dot(Ctx &c, const double *a, const double *b, int64_t n, double *out)
{
   const int g = vec_grid(n);
   pcg::dot_partial<1><<<g, pcg::RB, 0, c.stream>>>(n, 0, a, b, nullptr, c.d_part);
   LAGB_LAUNCH_CHECK();
   pcg::reduce_partials<1><<<1, pcg::RB, 0, c.stream>>>(g, c.d_part, c.d_tmp + 8);
   LAGB_LAUNCH_CHECK();
   int rc = allreduce_sum(c, c.d_tmp + 8, 1); if (rc) { return rc; }
   LAGB_CUDA(cudaMemcpyAsync(c.h_scal + 8, c.d_tmp + 8, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
   LAGB_CUDA(cudaStreamSynchronize(c.stream));
   *out = c.h_scal[8];
   return LAGB_OK;
}

int lagb_internal_energy(lagb_ctx *h, const double *d_e, double *h_out)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (!c.setup_done) { set_error("internal_energy: call lagb_setup_qdata0 first"); return LAGB_ERR_STATE; }
   KernelSet &ks = (c.variant == 1) ? c.ks_generic : c.ks;
   int rc = ks.mass_l2(c, d_e, c.d_lz); if (rc) { return rc; }                  // M_L2 e (element local)
   pcg::vec_fill<<<vec_grid(c.ndofs_l2), pcg::RB, 0, c.stream>>>(c.d_lr, 1.0, c.ndofs_l2);
   LAGB_LAUNCH_CHECK();
   return global_dot(c, c.d_lr, c.d_lz, c.ndofs_l2, h_out);
}

int lagb_kinetic_energy(lagb_ctx *h, const double *d_v, double *h_out)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (!c.setup_done) { set_error("kinetic_energy: call lagb_setup_qdata0 first"); return LAGB_ERR_STATE; }
   KernelSet &ks = (c.variant == 1) ? c.ks_generic : c.ks;
   const int64_t n = c.ndofs*c.dim;
   // scratch: the PCG residual vector (d_z / d_d2 stay zero between solves, see pcg_run_nc)
   LAGB_CUDA(cudaMemsetAsync(c.d_r, 0, sizeof(double)*n, c.stream));
   // rank-local element contributions only (no shared-dof sum): v^t (M_loc v) summed over ranks
   // counts every element once
   int rc = ks.mass_h1(c, c.dim, d_v, c.d_r, false); if (rc) { return rc; }
   double s = 0.0;
   rc = global_dot(c, d_v, c.d_r, n, &s); if (rc) { return rc; }
   *h_out = 0.5*s;
   return LAGB_OK;
}

int lagb_compute_density(lagb_ctx *h, const double *d_x, double *d_rho)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (!c.setup_done) { set_error("compute_density: call lagb_setup_qdata0 first"); return LAGB_ERR_STATE; }
   if (c.L1D < 1) { set_error("compute_density: needs a thermodynamic space"); return LAGB_ERR_INVALID; }
   double *wdet = nullptr;
   LAGB_CUDA(cudaMalloc((void**)&wdet, sizeof(double)*(size_t)c.NE*c.NQ));
   int rc = c.ks_generic.detj_w(c, d_x, wdet);
   if (!rc) { rc = density_project(c, wdet, d_rho); }
   cudaStreamSynchronize(c.stream);
   cudaFree(wdet);
   return rc;
}

int lagb_taylor_source(lagb_ctx *h, const double *d_x, double *d_esrc)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (c.dim != 2) { set_error("taylor_source: 2D only (reference laghos.cpp:638)"); return LAGB_ERR_INVALID; }
   return c.ks_generic.taylor(c, d_x, d_esrc);
}

double *lagb_qdata_ptr(lagb_ctx *h, int which)
{
   Ctx &c = h->c;
   switch (which)
   {
      case 0: return c.d_sJit; case 1: return c.d_rho0DetJ0w; case 2: return c.d_Jac0inv;
      case 3: return c.d_massD; case 4: return c.d_diag;
   }
   return nullptr;
}
double lagb_qdata_h0(const lagb_ctx *h) { return h->c.h0; }
int lagb_qdata_set_h0(lagb_ctx *h, double h0) { h->c.h0 = h0; return LAGB_OK; }

int lagb_dev_malloc(lagb_ctx *h, double **d_out, int64_t n)
{
   LAGB_ENTER(h);
   (void)h;
   LAGB_CUDA(cudaMalloc((void**)d_out, std::max<int64_t>(n, 1)*sizeof(double)));
   return LAGB_OK;
}
int lagb_dev_free(lagb_ctx *h, double *p) { (void)h; LAGB_CUDA(cudaFree(p)); return LAGB_OK; }
int lagb_memcpy_h2d(lagb_ctx *h, double *d_dst, const double *h_src, int64_t n)
{
   LAGB_ENTER(h);
   LAGB_CUDA(cudaMemcpyAsync(d_dst, h_src, sizeof(double)*n, cudaMemcpyHostToDevice, h->c.stream));
   LAGB_CUDA(cudaStreamSynchronize(h->c.stream));
   return LAGB_OK;
}
int lagb_memcpy_h2d_async(lagb_ctx *h, double *d_dst, const double *h_src, int64_t n)
{
   LAGB_ENTER(h);
   LAGB_CUDA(cudaMemcpyAsync(d_dst, h_src, sizeof(double)*n, cudaMemcpyHostToDevice, h->c.stream));
   return LAGB_OK;
}
int lagb_memcpy_d2h(lagb_ctx *h, double *h_dst, const double *d_src, int64_t n)
{
   LAGB_ENTER(h);
   LAGB_CUDA(cudaMemcpyAsync(h_dst, d_src, sizeof(double)*n, cudaMemcpyDeviceToHost, h->c.stream));
   LAGB_CUDA(cudaStreamSynchronize(h->c.stream));
   return LAGB_OK;
}
int lagb_memcpy_h2d_bg(lagb_ctx *h, double *d_dst, const double *h_src, int64_t n)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   LAGB_CUDA(cudaEventRecord(c.ev_compute, c.stream));
   LAGB_CUDA(cudaStreamWaitEvent(c.copy_stream, c.ev_compute, 0));
   LAGB_CUDA(cudaMemcpyAsync(d_dst, h_src, sizeof(double)*n, cudaMemcpyHostToDevice, c.copy_stream));
   LAGB_CUDA(cudaEventRecord(c.ev_copy, c.copy_stream));
   return LAGB_OK;
}
int lagb_memcpy_d2h_bg(lagb_ctx *h, double *h_dst, const double *d_src, int64_t n)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   LAGB_CUDA(cudaEventRecord(c.ev_compute, c.stream));
   LAGB_CUDA(cudaStreamWaitEvent(c.copy_stream, c.ev_compute, 0));
   LAGB_CUDA(cudaMemcpyAsync(h_dst, d_src, sizeof(double)*n, cudaMemcpyDeviceToHost, c.copy_stream));
   LAGB_CUDA(cudaEventRecord(c.ev_copy, c.copy_stream));
   return LAGB_OK;
}
int lagb_wait_copies(lagb_ctx *h)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   LAGB_CUDA(cudaStreamWaitEvent(c.stream, c.ev_copy, 0));
   return LAGB_OK;
}
int lagb_host_alloc_pinned(double **h_out, int64_t n)
{
   LAGB_CUDA(cudaMallocHost((void**)h_out, std::max<int64_t>(n, 1)*sizeof(double)));
   return LAGB_OK;
}
int lagb_host_free_pinned(double *p) { LAGB_CUDA(cudaFreeHost(p)); return LAGB_OK; }

int lagb_vec_fill(lagb_ctx *h, double *y, double a, int64_t n)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   pcg::vec_fill<<<vec_grid(n), pcg::RB, 0, c.stream>>>(y, a, n);
   LAGB_LAUNCH_CHECK();
   return LAGB_OK;
}
int lagb_vec_copy(lagb_ctx *h, double *y, const double *x, int64_t n)
{
   LAGB_ENTER(h);
   LAGB_CUDA(cudaMemcpyAsync(y, x, sizeof(double)*n, cudaMemcpyDeviceToDevice, h->c.stream));
   return LAGB_OK;
}
int lagb_vec_axpby(lagb_ctx *h, double *z, double a, const double *x, double b, const double *y, int64_t n)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   pcg::vec_axpby<<<vec_grid(n), pcg::RB, 0, c.stream>>>(z, a, x, b, y ? y : x, n);
   LAGB_LAUNCH_CHECK();
   return LAGB_OK;
}
int lagb_vec_dot(lagb_ctx *h, const double *x, const double *y, int64_t n, double *out)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   const int g = vec_grid(n);
   pcg::dot_partial<1><<<g, pcg::RB, 0, c.stream>>>(n, 0, x, y, nullptr, c.d_part);
   LAGB_LAUNCH_CHECK();
   pcg::reduce_partials<1><<<1, pcg::RB, 0, c.stream>>>(g, c.d_part, c.d_tmp + 8);
   LAGB_LAUNCH_CHECK();
   LAGB_CUDA(cudaMemcpyAsync(c.h_scal + 8, c.d_tmp + 8, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
   LAGB_CUDA(cudaStreamSynchronize(c.stream));
   *out = c.h_scal[8];
   return LAGB_OK;
}

int lagb_nccl_unique_id(uint8_t id_out[128])
{
   int rc = nccl_load(); if (rc) { return rc; }
   LAGB_NCCL(g_nccl.GetUniqueId(id_out));
   return LAGB_OK;
}

int lagb_ctx_comm_init(lagb_ctx *h, const uint8_t id[128], int rank, int nranks,
                       int nnbr, const int32_t *nbr_rank, const int32_t *exchange_phase,
                       const int32_t *nshared, const int32_t *const *h_shared_dofs,
                       const uint8_t *h_owner_mask)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   int rc = nccl_load(); if (rc) { return rc; }
   Uid128 uid; memcpy(uid.internal, id, 128);
   typedef int (*init_t)(void **, int, Uid128, int);
   LAGB_NCCL(((init_t)g_nccl.CommInitRankRaw)(&c.nccl_comm, nranks, uid, rank));
   c.rank = rank; c.nranks = nranks;
   c.nphases = 0;
   for (int k = 0; k < nnbr; k++)
   {
      Ctx::Nbr nb; nb.rank = nbr_rank[k]; nb.phase = exchange_phase[k]; nb.n = nshared[k];
      nb.d_idx = nullptr; nb.d_send = nullptr; nb.d_recv = nullptr;
      rc = dev_upload(&nb.d_idx, (const int*)h_shared_dofs[k], (size_t)nb.n); if (rc) { return rc; }
      rc = dev_alloc(&nb.d_send, (size_t)nb.n*3); if (rc) { return rc; }
      rc = dev_alloc(&nb.d_recv, (size_t)nb.n*3); if (rc) { return rc; }
      c.nbrs.push_back(nb);
      c.nphases = std::max(c.nphases, nb.phase + 1);
   }
   if (h_owner_mask) { rc = dev_upload(&c.d_own, (const unsigned char*)h_owner_mask, (size_t)c.ndofs); if (rc) { return rc; } }
   // all neighbours in phase 0: single-phase exchange with rank-ordered summation
   c.halo_single = (nnbr > 0 && c.nphases == 1 && nnbr < 255);
   if (c.halo_single)
   {
      std::vector<int> pack_idx, off(nnbr), cnt(nnbr);
      std::vector<unsigned char> nbk;
      int total = 0;
      for (int k = 0; k < nnbr; k++)
      {
         off[k] = total; cnt[k] = nshared[k];
         for (int j = 0; j < nshared[k]; j++) { pack_idx.push_back(h_shared_dofs[k][j]); nbk.push_back((unsigned char)k); }
         total += nshared[k];
      }
      // per shared dof: (rank, source) pairs, self = -1, sorted by rank
      std::vector<std::vector<std::pair<int,int>>> per((size_t)c.ndofs);
      for (int J = 0; J < total; J++) { per[pack_idx[J]].push_back({nbr_rank[nbk[J]], J}); }
      std::vector<int> u_dof, u_ptr(1, 0), u_src;
      for (int64_t i = 0; i < c.ndofs; i++)
      {
         if (per[i].empty()) { continue; }
         per[i].push_back({rank, -1});
         std::sort(per[i].begin(), per[i].end());
         u_dof.push_back((int)i);
         for (auto &pr : per[i]) { u_src.push_back(pr.second); }
         u_ptr.push_back((int)u_src.size());
      }
      c.halo_total = total; c.halo_nu = (int)u_dof.size(); c.h_nbr_off = off; c.h_nbr_n = cnt;
      rc = dev_upload(&c.d_pack_idx, pack_idx.data(), pack_idx.size()); if (rc) { return rc; }
      rc = dev_upload(&c.d_pack_nb, nbk.data(), nbk.size()); if (rc) { return rc; }
      rc = dev_upload(&c.d_nbr_off, off.data(), off.size()); if (rc) { return rc; }
      rc = dev_upload(&c.d_nbr_n, cnt.data(), cnt.size()); if (rc) { return rc; }
      rc = dev_upload(&c.d_u_dof, u_dof.data(), u_dof.size()); if (rc) { return rc; }
      rc = dev_upload(&c.d_u_ptr, u_ptr.data(), u_ptr.size()); if (rc) { return rc; }
      rc = dev_upload(&c.d_u_src, u_src.data(), u_src.size()); if (rc) { return rc; }
      rc = dev_alloc(&c.d_send_all, (size_t)total*3); if (rc) { return rc; }
      rc = dev_alloc(&c.d_recv_all, (size_t)total*3); if (rc) { return rc; }
   }
   return p2p_setup(c, nnbr, nbr_rank);
}

int lagb_allreduce_host(lagb_ctx *h, double *vals, int n, int op)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (c.nranks <= 1) { return LAGB_OK; }
   if (n > 8) { set_error("allreduce_host: n > 8"); return LAGB_ERR_INVALID; }
   const int nccl_op = (op == 0) ? NCCL_SUM : (op == 1) ? NCCL_MIN : 2 /* ncclMax */;
   memcpy(c.h_scal, vals, sizeof(double)*n);
   LAGB_CUDA(cudaMemcpyAsync(c.d_tmp + 8, c.h_scal, sizeof(double)*n, cudaMemcpyHostToDevice, c.stream));
   LAGB_NCCL(g_nccl.AllReduce(c.d_tmp + 8, c.d_tmp + 8, n, NCCL_F64, nccl_op, c.nccl_comm, c.stream));
   LAGB_CUDA(cudaMemcpyAsync(c.h_scal, c.d_tmp + 8, sizeof(double)*n, cudaMemcpyDeviceToHost, c.stream));
   LAGB_CUDA(cudaStreamSynchronize(c.stream));
   memcpy(vals, c.h_scal, sizeof(double)*n);
   return LAGB_OK;
}

int lagb_timing_get(lagb_ctx *h, lagb_timing *out)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   int rc = timer_resolve(c); if (rc) { return rc; }
   out->t_cgH1 = c.timer.acc[0]; out->t_cgL2 = c.timer.acc[1]; out->t_force = c.timer.acc[2]; out->t_qdata = c.timer.acc[3];
   out->H1iter = c.H1iter; out->L2iter = c.L2iter; out->quad_tstep = c.quad_tstep;
   return LAGB_OK;
}
int lagb_timing_reset(lagb_ctx *h)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   int rc = timer_resolve(c); if (rc) { return rc; }
   for (int w = 0; w < Timer::NT; w++) { c.timer.acc[w] = 0.0; }
   c.H1iter = c.L2iter = c.quad_tstep = 0; c.mass_launches = 0;
   return LAGB_OK;
}

int lagb_profile_mass(lagb_ctx *h, int enable) { h->c.profile_mass = enable != 0; return LAGB_OK; }
int lagb_profile_mass_get(lagb_ctx *h, double *seconds, int64_t *launches)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   int rc = timer_resolve(c); if (rc) { return rc; }
   if (seconds) { *seconds = c.timer.acc[4]; }
   if (launches) { *launches = c.mass_launches; }
   return LAGB_OK;
}
int lagb_stopwatch_start(lagb_ctx *h)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   int rc = timer_resolve(c); if (rc) { return rc; }
   c.timer.acc[5] = 0.0;
   return timer_begin(c, 5);
}
int lagb_stopwatch_stop(lagb_ctx *h, double *seconds)
{
   LAGB_ENTER(h);
   Ctx &c = h->c;
   if (c.timer.pending[5].empty()) { set_error("stopwatch_stop without start"); return LAGB_ERR_STATE; }
   int rc = timer_end(c, 5); if (rc) { return rc; }
   rc = timer_resolve(c); if (rc) { return rc; }
   if (seconds) { *seconds = c.timer.acc[5]; }
   return LAGB_OK;
}

} // extern "C"
