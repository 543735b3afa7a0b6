// Cartesian element partition of a rectilinear mesh over a process grid
// (SURVEY.md 8e): the B200 replacement of the reference's MPI domain decomposition
// (ParMesh from PartitionMPI / METIS, laghos.cpp:395-398, 481-483) for one 8-GPU
// NVSwitch box.  Each rank owns a box of elements; H1 dofs on box faces are
// duplicated on the sharing ranks.  This header computes, for one rank,
//   * its element box,
//   * the neighbour list with the shared scalar-dof indices in matching order and an
//     exchange phase per neighbour (phase = axis: three successive face exchanges sum
//     edge and corner dofs over all sharers, like P^t followed by P),
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
      for (int d = 0; d < dim; d++)
      {
         for (int side = 0; side < 2; side++)
         {
            const int nc = pc[d] + (side ? 1 : -1);
            if (nc < 0 || nc >= pgrid[d]) { continue; }
            int q[3] = {pc[0], pc[1], pc[2]}; q[d] = nc;
            Nbr nb; nb.rank = q[0] + pgrid[0]*(q[1] + pgrid[1]*q[2]); nb.phase = d;
            const int plane = side ? N1[d] - 1 : 0;
            for (int gz = 0; gz < N1[2]; gz++)
               for (int gy = 0; gy < N1[1]; gy++)
                  for (int gx = 0; gx < N1[0]; gx++)
                  {
                     const int g[3] = {gx, gy, gz};
                     if (g[d] != plane) { continue; }
                     const int id = gx + N1[0]*(gy + N1[1]*gz);
                     nb.dofs.push_back(id);
                     if (side == 0) { owner[id] = 0; }   // the lower neighbour owns the interface
                  }
            nbrs.push_back(std::move(nb));
         }
      }
   }
};

} // namespace lagb
