// Cartesian element partition of a rectilinear mesh over a process grid
// (SURVEY.md 8e): the B200 replacement of the reference's MPI domain decomposition
// (ParMesh from PartitionMPI / METIS, laghos.cpp:395-398, 481-483) for one 8-GPU
// NVSwitch box.  Each rank owns a box of elements; H1 dofs on box faces are
// duplicated on the sharing ranks.  This header computes, for one rank,
//   * its element box,
//   * the list of all ranks sharing dofs with it (face, edge and corner neighbours) with the
//     shared scalar-dof indices in matching (lexicographic lattice) order; one exchange phase,
//     contributions summed in ascending rank order (P^t followed by P in the reference),
//   * the owner mask used in inner products (each shared dof counted once).
#pragma once
#include <vector>
#include <cstdint>
#include <stdexcept>

namespace lagb {

struct Partition
{
   int dim = 3, rank = 0, nranks = 1;
   int pgrid[3] = {1, 1, 1}, pc[3] = {0, 0, 0};    // process grid and this rank's coordinates
   int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};       // element box
   struct Nbr { int rank, phase; std::vector<int> dofs; };
   std::vector<Nbr> nbrs;
   std::vector<unsigned char> owner;                // [local scalar H1 dofs]

   static void split(int n, int parts, int k, int &a, int &b)
   {
      a = (int)((long long)n*k/parts); b = (int)((long long)n*(k + 1)/parts);
   }

   // n: global elements per axis; ok: H1 order
   void build(int dim_, const int *n, const int *pgrid_, int rank_, int ok)
   {
      dim = dim_; rank = rank_;
      nranks = 1;
      for (int d = 0; d < 3; d++) { pgrid[d] = (d < dim) ? pgrid_[d] : 1; nranks *= pgrid[d]; }
      if (rank < 0 || rank >= nranks) { throw std::runtime_error("partition: bad rank"); }
      pc[0] = rank % pgrid[0]; pc[1] = (rank/pgrid[0]) % pgrid[1]; pc[2] = rank/(pgrid[0]*pgrid[1]);
      int N1[3] = {1, 1, 1};
      for (int d = 0; d < 3; d++)
      {
         if (d < dim)
         {
            if (pgrid[d] > n[d]) { throw std::runtime_error("partition: more ranks than elements along an axis"); }
            split(n[d], pgrid[d], pc[d], lo[d], hi[d]);
         }
         else { lo[d] = 0; hi[d] = 1; }
         N1[d] = (d < dim) ? (hi[d] - lo[d])*ok + 1 : 1;
      }
      const int64_t nd = (int64_t)N1[0]*N1[1]*N1[2];
      owner.assign((size_t)nd, 1);
      nbrs.clear();
      // All ranks that share at least one dof (faces, edges and corners: up to 26), one
      // exchange phase: every sharer sends its partial value to every other sharer and the
      // receiver sums the contributions in ascending rank order (the same order on every
      // sharer, so the copies of a shared dof stay bit-identical across ranks).
      for (int oz = -1; oz <= 1; oz++)
         for (int oy = -1; oy <= 1; oy++)
            for (int ox = -1; ox <= 1; ox++)
            {
               const int off[3] = {ox, oy, oz};
               if (ox == 0 && oy == 0 && oz == 0) { continue; }
               int q[3]; bool ok_n = true;
               for (int d = 0; d < 3; d++)
               {
                  q[d] = pc[d] + off[d];
                  if (d >= dim && off[d] != 0) { ok_n = false; }
                  if (q[d] < 0 || q[d] >= pgrid[d]) { ok_n = false; }
               }
               if (!ok_n) { continue; }
               Nbr nb; nb.rank = q[0] + pgrid[0]*(q[1] + pgrid[1]*q[2]); nb.phase = 0;
               int lo_i[3], hi_i[3];
               for (int d = 0; d < 3; d++)
               {
                  lo_i[d] = (off[d] == 1) ? N1[d] - 1 : 0;
                  hi_i[d] = (off[d] == -1) ? 0 : N1[d] - 1;
               }
               for (int gz = lo_i[2]; gz <= hi_i[2]; gz++)
                  for (int gy = lo_i[1]; gy <= hi_i[1]; gy++)
                     for (int gx = lo_i[0]; gx <= hi_i[0]; gx++) { nb.dofs.push_back(gx + N1[0]*(gy + N1[1]*gz)); }
               nbrs.push_back(std::move(nb));
            }
      // a dof on a lower face of the box belongs to a lower rank
      for (int gz = 0; gz < N1[2]; gz++)
         for (int gy = 0; gy < N1[1]; gy++)
            for (int gx = 0; gx < N1[0]; gx++)
            {
               const int g[3] = {gx, gy, gz};
               for (int d = 0; d < dim; d++) { if (g[d] == 0 && pc[d] > 0) { owner[gx + N1[0]*(gy + N1[1]*gz)] = 0; } }
            }
   }
};

} // namespace lagb
