// Primitives of the NVLink peer-memory exchanges (see device/p2p.cuh): buffer layout, system-scope
// loads / stores, flag wait, and the rank-ordered sum of NC scalars that the PCG finish kernels call.
#pragma once
#include "common.cuh"

namespace lagb {
namespace p2p {

constexpr int MAXR = 64;          // ranks
constexpr int SLOTW = 4;          // doubles per (parity, rank) scalar slot

struct Layout                     // byte offsets inside every rank's communication buffer
{
   size_t scal, sflag, hflag, halo[2], err;
   size_t bytes;
};
struct Dev                        // passed by value to the kernels
{
   char *peer[MAXR];              // mapped base address of every rank's buffer (peer[rank] = own)
   int rank, nranks;
   Layout lay;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
   asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
   unsigned long long v;
   asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ void st_relaxed_sys(double *p, double v)
{
   asm volatile("st.relaxed.sys.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p)
{
   double v;
   asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ bool wait_flag(const unsigned long long *p, unsigned long long seq, char *own_base, const Layout &lay)
{
   for (unsigned int it = 0; it < (1u << 28); it++)
   {
      if (ld_acquire_sys(p) >= seq) { return true; }
      if (it > 64) { __nanosleep(40); }
   }
   *reinterpret_cast<volatile int*>(own_base + lay.err) = 1;
   return false;
}

// Sum of `vals` (valid on thread 0; NC <= SLOTW) over all ranks in ascending rank order: publish into every rank's
// slot [parity][my rank], raise the flag, wait for every rank's flag, add.  Called by ALL threads of ONE CTA
// (blockDim.x >= nranks); the result is valid on thread 0.  `sh` needs NC doubles of shared memory.
template<int NC>
__device__ __forceinline__ void allreduce_cta(const Dev &d, const unsigned long long seq, double *vals, double *sh)
{
   if (threadIdx.x == 0) { for (int c = 0; c < NC; c++) { sh[c] = vals[c]; } }
   __syncthreads();
   const int par = (int)(seq & 1ull);
   const int t = threadIdx.x;
   if (t < d.nranks)
   {
      double *slot = reinterpret_cast<double*>(d.peer[t] + d.lay.scal) + ((size_t)par*d.nranks + d.rank)*SLOTW;
      for (int c = 0; c < NC; c++) { st_relaxed_sys(slot + c, sh[c]); }
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned long long*>(d.peer[t] + d.lay.sflag) + (size_t)par*d.nranks + d.rank, seq);
      wait_flag(reinterpret_cast<const unsigned long long*>(d.peer[d.rank] + d.lay.sflag) + (size_t)par*d.nranks + t, seq,
                d.peer[d.rank], d.lay);
   }
   __syncthreads();
   if (t == 0)
   {
      const double *slots = reinterpret_cast<const double*>(d.peer[d.rank] + d.lay.scal) + (size_t)par*d.nranks*SLOTW;
      for (int c = 0; c < NC; c++)
      {
         double s = 0.0;
         for (int r = 0; r < d.nranks; r++) { s += ld_relaxed_sys(slots + (size_t)r*SLOTW + c); }
         vals[c] = s;
      }
   }
}

} // namespace p2p
} // namespace lagb
