// Host-side schedule of the "brick" H1 mass apply (device/mass3d_brick.cuh).
//
// The reference's E^t (MFEM ElementRestriction::MultTranspose behind
// MassPAOperator::Mult, laghos_assembly.cpp:117-121) sums the element contributions of a
// shared dof in a fixed order.  The first version of this library scattered with
// red.global.add.f64 (order not fixed, every dof gathered / added ~2.4 times).  The schedule
// built here removes both problems without giving up the general gather map:
//
//   * elements are grouped into BATCHES (one CTA each).  On a structured element grid
//     (grid hint nx,ny,nz: elements lexicographic, x fastest) a batch is a BX x BY x BZ brick,
//     otherwise NB consecutive elements;
//   * per batch: the sorted list of UNIQUE dofs it touches (`uid`), the element-local ->
//     unique-slot table (`lidx`) and its inverse in CSR form (`uoff`, `upos`), so the CTA
//     loads every dof once (coalesced along lattice rows), and sums the contributions of its
//     own elements to a dof in a fixed order inside shared memory;
//   * batches that share a dof get different COLOURS; one kernel launch per colour, in
//     colour order.  Within a launch no two CTAs touch the same dof, so the output is
//     written with plain stores: the lowest-coloured batch of a dof (flag bit 31 of `uid`,
//     "first writer") stores, every later one does load-add-store.  No atomics, no zero
//     fill of the output, and the summation order is fixed (colour order): deterministic;
//   * identical index tables (all interior bricks of a Cartesian mesh) are stored once.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace lagb {

struct BatchPlan
{
   int NB = 0, ND = 0, UP = 0;          // elements per batch, dofs per element, padded unique capacity
   int nbatch = 0, ncolors = 0, ntab = 0, umax = 0;
   int brick[3] = {0, 0, 0};            // brick extents in elements (0: unstructured batches)
   std::vector<int> color_begin;        // [ncolors+1] batch ranges, batches sorted by colour
   std::vector<int> elem;               // [nbatch*NB] element ids, -1 = padding
   std::vector<int> nuniq;              // [nbatch]
   std::vector<uint32_t> uid;           // [nbatch*UP] global scalar dof | (first writer << 31)
   std::vector<int> tab;                // [nbatch] index table id
   std::vector<uint16_t> lidx;          // [ntab*NB*ND] element-local dof -> unique slot
   std::vector<uint16_t> uoff;          // [ntab*(UP+1)] CSR offsets into upos
   std::vector<uint16_t> upos;          // [ntab*NB*ND] E positions (e_loc*ND + i) sorted by unique slot
   int64_t n_first = 0;                 // number of (dof) first-writer entries == number of touched dofs
   static constexpr int MAXDEP = 32;
   std::vector<int> deps;               // [nbatch*MAXDEP] earlier (lower-coloured) batches sharing a dof, -1 padded
   bool deps_ok = false;                // false: some batch has more than MAXDEP such neighbours

   static void brick_shape(int NB, const int grid[3], int b[3])
   {
      // split NB (power of two) over the axes, x first (longest rows = best coalescing), never
      // exceeding the grid extent
      b[0] = b[1] = b[2] = 1;
      int rem = NB, ax = 0, stuck = 0;
      while (rem > 1 && stuck < 3)
      {
         if (b[ax]*2 <= std::max(1, grid[ax])) { b[ax] *= 2; rem /= 2; stuck = 0; }
         else { stuck++; }
         ax = (ax + 1) % 3;
      }
   }

   // map: [NE*ND] scalar dof of element-local dof i (lexicographic).  grid: structured hint or {0,0,0}.
   // shape: requested brick extents (0 = choose from NB).
   int build(const int *map, int NE, int ND_, int64_t ndofs, const int grid[3], int NB_, const int shape[3], std::string &err)
   {
      NB = NB_; ND = ND_;
      if ((size_t)NB*ND > 65535) { err = "batch plan: NB*ND exceeds 16-bit positions"; return 1; }
      const bool structured = grid[0] > 0 && (int64_t)grid[0]*grid[1]*grid[2] == NE;
      std::vector<std::vector<int>> bel;     // elements of each batch
      std::vector<int> bcolor;
      if (structured)
      {
         if (shape && shape[0] > 0) { brick[0] = shape[0]; brick[1] = shape[1]; brick[2] = shape[2]; }
         else { brick_shape(NB, grid, brick); }
         if (brick[0]*brick[1]*brick[2] > NB) { err = "batch plan: brick larger than NB"; return 1; }
         const int nbk[3] = {(grid[0] + brick[0] - 1)/brick[0], (grid[1] + brick[1] - 1)/brick[1], (grid[2] + brick[2] - 1)/brick[2]};
         for (int kz = 0; kz < nbk[2]; kz++)
            for (int ky = 0; ky < nbk[1]; ky++)
               for (int kx = 0; kx < nbk[0]; kx++)
               {
                  std::vector<int> el;
                  for (int z = kz*brick[2]; z < std::min(grid[2], (kz + 1)*brick[2]); z++)
                     for (int y = ky*brick[1]; y < std::min(grid[1], (ky + 1)*brick[1]); y++)
                        for (int x = kx*brick[0]; x < std::min(grid[0], (kx + 1)*brick[0]); x++)
                        { el.push_back(x + grid[0]*(y + grid[1]*z)); }
                  bel.push_back(el);
                  bcolor.push_back((kx & 1) | ((ky & 1) << 1) | ((kz & 1) << 2));
               }
      }
      else
      {
         brick[0] = brick[1] = brick[2] = 0;
         for (int e0 = 0; e0 < NE; e0 += NB)
         {
            std::vector<int> el;
            for (int e = e0; e < std::min(NE, e0 + NB); e++) { el.push_back(e); }
            bel.push_back(el);
         }
         bcolor.assign(bel.size(), -1);
      }
      nbatch = (int)bel.size();
      // unique dofs per batch
      std::vector<std::vector<int>> buniq(nbatch);
      umax = 0;
      for (int b = 0; b < nbatch; b++)
      {
         auto &u = buniq[b];
         for (int e : bel[b]) { u.insert(u.end(), map + (size_t)e*ND, map + (size_t)(e + 1)*ND); }
         std::sort(u.begin(), u.end());
         u.erase(std::unique(u.begin(), u.end()), u.end());
         umax = std::max(umax, (int)u.size());
      }
      if (umax > 65535) { err = "batch plan: too many unique dofs per batch"; return 1; }
      if (!structured)
      {
         // greedy colouring of the batch conflict graph (batches sharing a dof)
         std::vector<int> head((size_t)ndofs + 1, 0);
         for (int b = 0; b < nbatch; b++) { for (int d : buniq[b]) { head[d + 1]++; } }
         for (int64_t i = 0; i < ndofs; i++) { head[i + 1] += head[i]; }
         std::vector<int> d2b(head[ndofs]), fill(head.begin(), head.end() - 1);
         for (int b = 0; b < nbatch; b++) { for (int d : buniq[b]) { d2b[fill[d]++] = b; } }
         std::vector<char> used;
         for (int b = 0; b < nbatch; b++)
         {
            used.assign(64, 0);
            for (int d : buniq[b])
               for (int p = head[d]; p < head[d + 1]; p++)
               {
                  const int c = bcolor[d2b[p]];
                  if (c >= 0) { if (c >= (int)used.size()) { used.resize(c + 1, 0); } used[c] = 1; }
               }
            int c = 0; while (c < (int)used.size() && used[c]) { c++; }
            bcolor[b] = c;
         }
      }
      // compress colours, sort batches by (colour, original order)
      {
         std::vector<int> cs(bcolor); std::sort(cs.begin(), cs.end()); cs.erase(std::unique(cs.begin(), cs.end()), cs.end());
         ncolors = (int)cs.size();
         for (int &c : bcolor) { c = (int)(std::lower_bound(cs.begin(), cs.end(), c) - cs.begin()); }
      }
      std::vector<int> order(nbatch);
      for (int b = 0; b < nbatch; b++) { order[b] = b; }
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return bcolor[a] < bcolor[b]; });
      color_begin.assign(ncolors + 1, 0);
      for (int b = 0; b < nbatch; b++) { color_begin[bcolor[b] + 1]++; }
      for (int c = 0; c < ncolors; c++) { color_begin[c + 1] += color_begin[c]; }
      // first writer of a dof = its lowest-coloured batch
      std::vector<int> mincol((size_t)ndofs, 1 << 30);
      for (int b = 0; b < nbatch; b++) { for (int d : buniq[b]) { mincol[d] = std::min(mincol[d], bcolor[b]); } }
      UP = ((umax + 31)/32)*32;
      elem.assign((size_t)nbatch*NB, -1); nuniq.assign(nbatch, 0); uid.assign((size_t)nbatch*UP, 0u); tab.assign(nbatch, 0);
      std::map<std::string, int> tabs;
      lidx.clear(); uoff.clear(); upos.clear(); ntab = 0; n_first = 0;
      std::vector<uint16_t> li((size_t)NB*ND), uo((size_t)UP + 1), up((size_t)NB*ND);
      for (int k = 0; k < nbatch; k++)
      {
         const int b = order[k];
         const auto &u = buniq[b];
         nuniq[k] = (int)u.size();
         for (size_t j = 0; j < bel[b].size(); j++) { elem[(size_t)k*NB + j] = bel[b][j]; }
         for (size_t j = 0; j < u.size(); j++)
         {
            const bool first = (mincol[u[j]] == bcolor[b]);
            n_first += first;
            uid[(size_t)k*UP + j] = (uint32_t)u[j] | (first ? 0x80000000u : 0u);
         }
         // padding slots repeat the last dof without the first-writer flag (never stored: slot >= nuniq)
         for (int j = (int)u.size(); j < UP; j++) { uid[(size_t)k*UP + j] = u.empty() ? 0u : (uint32_t)u.back(); }
         std::fill(li.begin(), li.end(), (uint16_t)0);
         std::vector<int> cnt(UP + 1, 0);
         for (size_t j = 0; j < bel[b].size(); j++)
            for (int i = 0; i < ND; i++)
            {
               const int d = map[(size_t)bel[b][j]*ND + i];
               const int s = (int)(std::lower_bound(u.begin(), u.end(), d) - u.begin());
               li[j*ND + i] = (uint16_t)s;
               cnt[s + 1]++;
            }
         for (int s = 0; s < UP; s++) { cnt[s + 1] += cnt[s]; }
         for (int s = 0; s <= UP; s++) { uo[s] = (uint16_t)cnt[s]; }
         std::vector<int> fillp(cnt.begin(), cnt.end() - 1);
         std::fill(up.begin(), up.end(), (uint16_t)0);
         for (size_t j = 0; j < bel[b].size(); j++)
            for (int i = 0; i < ND; i++) { up[fillp[li[j*ND + i]]++] = (uint16_t)(j*ND + i); }
         std::string key((const char*)li.data(), li.size()*2);
         key.append((const char*)uo.data(), uo.size()*2);
         key.push_back((char)bel[b].size());
         auto it = tabs.find(key);
         if (it == tabs.end())
         {
            it = tabs.emplace(key, ntab++).first;
            lidx.insert(lidx.end(), li.begin(), li.end());
            uoff.insert(uoff.end(), uo.begin(), uo.end());
            upos.insert(upos.end(), up.begin(), up.end());
         }
         tab[k] = it->second;
      }
      // dependency lists of the single-launch dataflow kernel (device/mass3d_brick3.cuh): for the batch at
      // schedule position k, the positions k' < k that share a dof with it (same colour never shares)
      {
         std::vector<int> pos(nbatch);
         for (int k = 0; k < nbatch; k++) { pos[order[k]] = k; }
         std::vector<int64_t> head((size_t)ndofs + 1, 0);
         for (int b = 0; b < nbatch; b++) { for (int d : buniq[b]) { head[d + 1]++; } }
         for (int64_t i = 0; i < ndofs; i++) { head[i + 1] += head[i]; }
         std::vector<int> d2k((size_t)head[ndofs]);
         std::vector<int64_t> fill(head.begin(), head.end() - 1);
         for (int b = 0; b < nbatch; b++) { for (int d : buniq[b]) { d2k[fill[d]++] = pos[b]; } }
         deps.assign((size_t)nbatch*MAXDEP, -1);
         deps_ok = true;
         std::vector<int> tmp;
         for (int k = 0; k < nbatch; k++)
         {
            tmp.clear();
            for (int d : buniq[order[k]])
            {
               if (head[d + 1] - head[d] < 2) { continue; }
               for (int64_t p = head[d]; p < head[d + 1]; p++) { if (d2k[p] < k) { tmp.push_back(d2k[p]); } }
            }
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            if ((int)tmp.size() > MAXDEP) { deps_ok = false; break; }
            for (size_t j = 0; j < tmp.size(); j++) { deps[(size_t)k*MAXDEP + j] = tmp[j]; }
         }
      }
      return 0;
   }

   // Invariants the kernels rely on (tests/test_host_setup.py through lagb_host_batch_plan_check).
   int self_check(const int *map, int NE, int64_t ndofs, std::string &err) const
   {
      std::vector<int> seen(NE, 0);
      std::vector<int> first((size_t)ndofs, 0), touched((size_t)ndofs, 0);
      for (int c = 0; c < ncolors; c++)
      {
         std::vector<int> owner((size_t)ndofs, -1);
         for (int k = color_begin[c]; k < color_begin[c + 1]; k++)
         {
            const int t = tab[k];
            const uint16_t *li = &lidx[(size_t)t*NB*ND], *uo = &uoff[(size_t)t*(UP + 1)], *up = &upos[(size_t)t*NB*ND];
            int nel = 0;
            for (int j = 0; j < NB; j++)
            {
               const int e = elem[(size_t)k*NB + j];
               if (e < 0) { continue; }
               if (j != nel) { err = "padding inside a batch"; return 1; }
               nel++;
               if (e >= NE || seen[e]++) { err = "element missing or repeated"; return 1; }
               for (int i = 0; i < ND; i++)
               {
                  const int s = li[j*ND + i];
                  if (s >= nuniq[k] || (int)(uid[(size_t)k*UP + s] & 0x7fffffffu) != map[(size_t)e*ND + i]) { err = "lidx/uid mismatch"; return 1; }
               }
            }
            if (uo[nuniq[k]] != nel*ND) { err = "CSR size"; return 1; }
            for (int s = 0; s < nuniq[k]; s++)
            {
               const uint32_t w = uid[(size_t)k*UP + s];
               const int d = (int)(w & 0x7fffffffu);
               if (s > 0 && d <= (int)(uid[(size_t)k*UP + s - 1] & 0x7fffffffu)) { err = "uid not sorted"; return 1; }
               if (owner[d] >= 0) { err = "two batches of one colour share a dof"; return 1; }
               owner[d] = k;
               if (w >> 31) { if (touched[d]) { err = "first writer is not the first"; return 1; } first[d]++; }
               else if (!touched[d]) { err = "dof updated before its first writer"; return 1; }
               if (uo[s + 1] <= uo[s]) { err = "empty CSR row"; return 1; }
               for (int p = uo[s]; p < uo[s + 1]; p++)
               {
                  if (li[up[p]] != s) { err = "upos/lidx mismatch"; return 1; }
                  if (p > uo[s] && up[p] <= up[p - 1]) { err = "upos not sorted"; return 1; }
               }
            }
            for (int s = 0; s < nuniq[k]; s++) { touched[uid[(size_t)k*UP + s] & 0x7fffffffu] = 1; }
         }
      }
      for (int e = 0; e < NE; e++) { if (seen[e] != 1) { err = "element not scheduled"; return 1; } }
      for (int64_t d = 0; d < ndofs; d++) { if (touched[d] && first[d] != 1) { err = "first-writer count"; return 1; } }
      return 0;
   }
};

} // namespace lagb
