// Internal context of the C ABI (include/laghos_b200.h).
#pragma once
#include "../../include/laghos_b200.h"
#include "device/pcg.cuh"
#include "device/p2p.cuh"
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>

namespace lagb {

void set_error(const std::string &msg);
extern int64_t g_launch_count;

#define LAGB_CUDA(call) do { cudaError_t err__ = (call); if (err__ != cudaSuccess) { \
   lagb::set_error(std::string(#call) + ": " + cudaGetErrorString(err__)); return LAGB_ERR_CUDA; } } while (0)
#define LAGB_LAUNCH_CHECK() do { lagb::g_launch_count++; cudaError_t err__ = cudaGetLastError(); if (err__ != cudaSuccess) { \
   lagb::set_error(std::string("kernel launch: ") + cudaGetErrorString(err__)); return LAGB_ERR_CUDA; } } while (0)

struct Ctx;

// Kernel launch with (pdl = true) or without the programmatic-stream-serialisation attribute: the kernel may be
// scheduled while its predecessor in the stream drains; it calls pdl_wait() (device/common.cuh) before its first
// dependent access.
template<typename... KArgs, typename... Args>
inline cudaError_t launch_k(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
   cudaLaunchConfig_t cfg = {};
   cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
   cudaLaunchAttribute at[1];
   at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
   at[0].val.programmaticStreamSerializationAllowed = 1;
   cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
   return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define LAGB_LAUNCH_K(c, kern, grid, block, smem, ...) do { lagb::g_launch_count++; \
   cudaError_t err__ = lagb::launch_k((c).tune[12] == 0, kern, dim3(grid), dim3(block), smem, (c).stream, __VA_ARGS__); \
   if (err__ != cudaSuccess) { lagb::set_error(std::string("kernel launch: ") + cudaGetErrorString(err__)); return LAGB_ERR_CUDA; } } while (0)

// input of the brick mass apply (device/mass3d_brick.cuh): plain x, or the fused PCG direction
// update d_new = M^-1 r + beta d_old (x == nullptr)
struct MassBrickIn
{
   const double *x = nullptr;
   const double *r = nullptr, *dold = nullptr; double *dnew = nullptr;
   int comp0 = 0;
};

// device copy of a host BatchPlan (host/batch_plan.hpp)
struct DevPlan
{
   int NB = 0, UP = 0, nbatch = 0, ncolors = 0, ntab = 0;
   std::vector<int> color_begin;
   int *belem = nullptr, *bnuniq = nullptr, *btab = nullptr; uint32_t *buid = nullptr;
   uint16_t *lidx = nullptr, *uoff = nullptr, *upos = nullptr;
   uint16_t *ucon = nullptr;          // [ntab][UP][8] fixed-width contribution table (plane slots), nullptr if a dof has > 8
   int brick[3] = {0, 0, 0};
   // single-launch dataflow kernel (device/mass3d_brick3.cuh)
   int *bmeta = nullptr, *deps = nullptr, *flags = nullptr, *work_ctr = nullptr; int epoch = 0;
};

// per-(DIM,D1D,Q1D) launchers
struct KernelSet
{
   int (*mass_h1)(Ctx&, int nc, const double *x, double *y, bool with_den) = nullptr; // y += M x (nc comps, byNODES stride)
   // y = M x without atomics (coloured brick schedule); nullptr where not instantiated
   int (*mass_brick)(Ctx&, int nc, const MassBrickIn &in, double *y, bool with_den) = nullptr;
   int (*mass_diag)(Ctx&, double *diag) = nullptr;
   int (*mass_l2)(Ctx&, const double *x, double *y) = nullptr;
   int (*force_mult)(Ctx&, const double *e, double *v) = nullptr;       // v += F e
   int (*force_mult_t)(Ctx&, const double *v, double *e) = nullptr;
   int (*qupdate)(Ctx&, const double *S, const QPointParams &prm) = nullptr; // writes dt_part, sets dt_nblocks
   int (*rho0detj0)(Ctx&, const double *x0, const double *rho0_gf, const double *rho0_q, double *elem_vol) = nullptr;
   int (*taylor)(Ctx&, const double *x, double *esrc) = nullptr;
   int (*detj_w)(Ctx&, const double *x, double *out) = nullptr;     // w detJ on the current mesh, [NE*NQ]
   bool tuned_mass = false;
};

struct Timer
{
   static constexpr int NT = 6;
   std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending[NT];
   std::vector<cudaEvent_t> pool;
   // cgH1, cgL2, force, qdata, [4] = H1 mass-apply kernel alone (lagb_profile_mass), [5] = stopwatch
   double acc[NT] = {0, 0, 0, 0, 0, 0};
   bool open[NT] = {false, false, false, false, false, false};   // newest interval has no end event yet
};

struct NcclApi;

struct Ctx
{
   int dim = 0, NE = 0, D1D = 0, L1D = 0, Q1D = 0, ND = 0, NL = 0, NQ = 0;
   int64_t ndofs = 0, ndofs_l2 = 0;
   int use_visc = 1, use_vort = 0, variant = 0, device = 0;
   cudaStream_t stream = nullptr;
   cudaStream_t copy_stream = nullptr;               // background state transfers (lagb_memcpy_*_bg)
   cudaEvent_t ev_compute = nullptr, ev_copy = nullptr;
   cudaStream_t aux_stream = nullptr;                // clears the PCG's idle result buffer behind the iteration
   cudaEvent_t ev_zuse = nullptr, ev_zclr[2] = {nullptr, nullptr}; bool z_pending[2] = {false, false}; bool z_clean = true;
   KernelSet ks, ks_generic;
   std::vector<unsigned char> tab_blob;   // DevTables<D1D,Q1D> bytes
   // device arrays
   int *d_map = nullptr; int *d_ess[3] = {nullptr, nullptr, nullptr}; int ness[3] = {0, 0, 0};
   double *d_qweights = nullptr, *d_inv_qweights = nullptr, *d_gamma = nullptr;
   double *d_sJit = nullptr, *d_rho0DetJ0w = nullptr, *d_Jac0inv = nullptr, *d_massD = nullptr;
   double *d_diag = nullptr, *d_dinv = nullptr;      // [ndofs] mass diagonal and its inverse
   unsigned char *d_essmask = nullptr;               // [ndofs] bit c: essential for component c
   double *d_r = nullptr, *d_d = nullptr, *d_z = nullptr;      // [dim*ndofs]
   double *d_d2 = nullptr;                                     // second search-direction buffer (fused brick PCG)
   int elem_grid[3] = {0, 0, 0};                               // structured element grid hint (0 = none)
   std::vector<int> h_map;                                     // host copy of the gather map (schedules are built lazily)
   std::map<const void*, int> occ_cache;                       // resident CTAs per SM of the persistent kernels
   int num_sms = 148;
   std::map<const void*, size_t> smem_optin;                   // kernels whose dynamic shared memory opt-in is set on this device
   std::map<int, DevPlan> plans;                               // brick schedules by (NB | shape key)
   double *d_lr = nullptr, *d_ld = nullptr, *d_lz = nullptr;   // [ndofs_l2]
   // element inverses of the L2 mass matrix (device/l2solve.cuh), built on the first energy solve
   double *d_l2inv = nullptr, *d_BL = nullptr; int l2inv_state = 0;   // 0 not tried, 1 ready, -1 unavailable (memory)
   double *d_part = nullptr; int part_cap = 0;                 // reduction partials
   unsigned int *d_grp_ctr = nullptr; int grp_cap = 0; size_t den_off = 0;   // arrival ticket of the multi-CTA finish kernels
   double *d_fin = nullptr;                                    // their per-CTA chunk sums
   double *d_tmp = nullptr;                                    // [8] reduced scalars
   double *d_dt = nullptr;                                     // [1]
   double *d_elem_vol = nullptr;
   pcg::State *d_state = nullptr;
   pcg::State *h_state = nullptr;   // pinned
   double *h_scal = nullptr;        // pinned [8]
   unsigned char *d_own = nullptr;  // owner mask (multi-rank), else nullptr
   double h0 = 0.0;
   bool setup_done = false;
   int dt_nblocks = 0;
   // lagb_tune_set: [0] legacy mass3d NC=3 variant, [1] force, [2] qupdate, [3] legacy mass3d NC=1, [4] brick launch
   // variant, [5] 1 = no programmatic dependent launch, [6] mass path (0 default, 1 legacy atomic, 2 brick v1, 3 brick v2),
   // [7] brick shape (0 cube-like, 1 x-long, 2 x-pencil), [8] 1 = first PCG vector kernels (update_xr/update_d),
   // [9] 1 = plain (not fused) PCG on the multi-launch brick kernels, [10] 1 = energy solve by the reference's CG
   // instead of the element inverses, [11] 1 = NCCL send/recv + all-reduce instead of the peer-memory exchanges,
   // [12] 1 = plain launches (no programmatic dependent launch) in the PCG iteration,
   // [13] 1 = zero fill of A d inside the direction kernel (one result buffer) instead of the alternating buffers,
   // [15] 1 = fixed 148*16 grid for the PCG vector kernels instead of one wave (occupancy x SMs)
   int tune[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
   int predicted_iters = 0;
   // timing
   Timer timer; int64_t H1iter = 0, L2iter = 0, quad_tstep = 0;
   bool profile_mass = false; int64_t mass_launches = 0;
   // multi-rank
   int rank = 0, nranks = 1; void *nccl_comm = nullptr;
   struct Nbr { int rank, phase, n; int *d_idx; double *d_send, *d_recv; };
   std::vector<Nbr> nbrs; int nphases = 0;
   // single-phase exchange (all sharers at once, contributions summed in ascending rank order)
   bool halo_single = false;
   int halo_total = 0, halo_nu = 0;                 // concatenated shared entries, distinct shared dofs
   std::vector<int> h_nbr_off, h_nbr_n;             // per neighbour: offset / count in the concatenation
   int *d_pack_idx = nullptr, *d_nbr_off = nullptr, *d_nbr_n = nullptr, *d_u_dof = nullptr, *d_u_ptr = nullptr, *d_u_src = nullptr;
   unsigned char *d_pack_nb = nullptr;
   double *d_send_all = nullptr, *d_recv_all = nullptr;
   int64_t ne_global = 0;
   // NVLink peer-memory exchanges (device/p2p.cuh): every rank's communication buffer mapped through cudaIpc
   bool p2p_on = false;
   p2p::Dev p2p_dev; p2p::Dev *d_p2p_dev = nullptr;  // host copy (kernel parameter) and device copy (finish kernels)
   char *p2p_base = nullptr;                        // own buffer (cudaMalloc)
   std::vector<void*> p2p_opened;                   // peers' mappings (cudaIpcCloseMemHandle on destroy)
   unsigned long long p2p_scal_seq = 0, p2p_halo_seq = 0;
   int *d_nbr_roff = nullptr, *d_nbr_rank = nullptr; unsigned int *d_pack_done = nullptr;
};

// helpers implemented in capi.cu
int vec_grid(int64_t n);
int timer_begin(Ctx &c, int which);
int timer_end(Ctx &c, int which);
int halo_sum(Ctx &c, double *v, int nc);              // sum shared dofs across ranks (no-op for 1 rank)
int allreduce_sum(Ctx &c, double *d_vals, int n);     // in-stream
int allreduce_min(Ctx &c, double *d_vals, int n);

int l2_direct_solve(Ctx &c, const double *b, double *x, bool *done);
int density_project(Ctx &c, const double *wdet, double *rho);        // kernels_l2.cu: ComputeDensity's element solves   // kernels_l2.cu; *done = false: not available
int get_plan(Ctx &c, int NB, DevPlan **out);    // brick schedule for NB elements per batch and the shape c.tune[7] (built on first use)

KernelSet make_generic_kernels(int dim, int D1D, int Q1D);
bool add_tuned_kernels(KernelSet &ks, int dim, int D1D, int Q1D);

} // namespace lagb
