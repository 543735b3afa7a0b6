// Fused reduce / exchange kernels over NVLink peer memory.
//
// The reference's parallel run needs two exchanges per CG iteration: the sum of the shared-dof partial
// values (P^t / P of the RAP operator, laghos_assembly.cpp:95) and the MPI_Allreduce of the inner products
// (MFEM CGSolver::Dot; laghos_solver.cpp:388 call site).  Through NCCL each one is a separate launch with
// 15-40 us of latency on 24 bytes / a few hundred KB, strictly serial inside the iteration.  Here every rank
// maps every other rank's communication buffer (cudaIpc handles exchanged once) and the exchanges are plain
// kernels of the iteration's stream:
//   p2p_allreduce   the last-stage reduction of the per-CTA partials, the publication of the NC sums into
//                   every rank's slot array (st.relaxed.sys + fence + flag with st.release.sys), the wait for
//                   the other ranks' flags and the fixed-rank-order sum, in ONE single-CTA launch;
//   halo_pack_p2p   packs the shared-dof values straight into the neighbours' receive areas, the last CTA
//                   to finish raises this rank's flag at every neighbour;
//   halo_combine_p2p waits for the neighbours' flags, then sums own and received values in ascending rank
//                   order (bit-identical on every sharer), as halo_combine does for the NCCL path.
// Slots, receive areas and flags are double-buffered by the parity of a sequence number that advances
// identically on all ranks (same stream of operations everywhere): a rank can be at most one exchange ahead
// of a peer, because finishing exchange k+1 needs the peer's publication k+1, which the peer's stream orders
// after its reads of exchange k.
// Spin loops give up after ~2^28 polls and set an error word instead of hanging the device.
#pragma once
#include "pcg.cuh"
#include "p2p_prims.cuh"

namespace lagb {
namespace p2p {

// out[c] = sum over ranks (ascending) of sum_b part[b*NC + c]
template<int NC>
__global__ void __launch_bounds__(pcg::FB)
p2p_allreduce(const Dev d, const unsigned long long seq, const int nblocks, const double *__restrict__ part,
              double *__restrict__ out)
{
   pdl_launch(); pdl_wait();
   __shared__ double sh[32];
   __shared__ double mine[NC];
   double tmp[NC];
   pcg::reduce_to_thread0<NC>(part, nblocks, tmp, sh);
   allreduce_cta<NC>(d, seq, tmp, mine);
   if (threadIdx.x == 0) { for (int c = 0; c < NC; c++) { out[c] = tmp[c]; } }
}

// pack every (neighbour, shared dof) entry into the neighbour's receive area; message layout per neighbour k:
// [c][j], base offset nc*roff[k] doubles (roff: where this rank's message starts in THAT rank's area)
static __global__ void halo_pack_p2p(const Dev d, const unsigned long long seq, int total, int nc, int64_t cstride,
                              const int *__restrict__ idx, const unsigned char *__restrict__ nbk,
                              const int *__restrict__ off, const int *__restrict__ cnt, const int *__restrict__ roff,
                              const int *__restrict__ nbr_rank, int nnbr, const double *__restrict__ v,
                              unsigned int *__restrict__ done)
{
   pdl_launch(); pdl_wait();
   const int par = (int)(seq & 1ull);
   for (int J = blockIdx.x*blockDim.x + threadIdx.x; J < total; J += gridDim.x*blockDim.x)
   {
      const int k = nbk[J], o = off[k], n = cnt[k], j = J - o;
      const int id = idx[J];
      double *dst = reinterpret_cast<double*>(d.peer[nbr_rank[k]] + d.lay.halo[par]) + (size_t)nc*roff[k] + j;
      for (int c = 0; c < nc; c++) { st_relaxed_sys(dst + (size_t)c*n, v[id + c*cstride]); }
   }
   __threadfence_system();
   __syncthreads();
   __shared__ bool last;
   if (threadIdx.x == 0) { last = (atomicAdd(done, 1u) == gridDim.x - 1); }
   __syncthreads();
   if (last)
   {
      __threadfence_system();
      for (int k = threadIdx.x; k < nnbr; k += blockDim.x)
      {
         st_release_sys(reinterpret_cast<unsigned long long*>(d.peer[nbr_rank[k]] + d.lay.hflag) + (size_t)par*d.nranks + d.rank, seq);
      }
      if (threadIdx.x == 0) { *done = 0u; }
   }
}

// v[dof] = sum over the sharers of dof (own value and received values) in ascending rank order
static __global__ void halo_combine_p2p(const Dev d, const unsigned long long seq, int nu, int nc, int64_t cstride,
                                 const int *__restrict__ u_dof, const int *__restrict__ u_ptr, const int *__restrict__ u_src,
                                 const unsigned char *__restrict__ nbk, const int *__restrict__ off, const int *__restrict__ cnt,
                                 const int *__restrict__ nbr_rank, int nnbr, double *__restrict__ v)
{
   pdl_launch(); pdl_wait();
   const int par = (int)(seq & 1ull);
   for (int k = threadIdx.x; k < nnbr; k += blockDim.x)
   {
      wait_flag(reinterpret_cast<const unsigned long long*>(d.peer[d.rank] + d.lay.hflag) + (size_t)par*d.nranks + nbr_rank[k], seq,
                d.peer[d.rank], d.lay);
   }
   __syncthreads();
   const double *recv = reinterpret_cast<const double*>(d.peer[d.rank] + d.lay.halo[par]);
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
               acc += __ldcg(recv + (size_t)nc*o + (size_t)c*n + (J - o));     // written by a peer: never through L1
            }
         }
         v[dof + c*cstride] = acc;
      }
   }
}

} // namespace p2p
} // namespace lagb
