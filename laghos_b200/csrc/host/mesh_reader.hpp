// Reader for the reference's mesh files (MFEM mesh v1.0, reference data/*.mesh), restricted to what
// the hot path's host setup supports: rectilinear quad / hex meshes whose boundary attributes follow
// the reference's convention attribute k <-> faces of constant x_{k-1} (laghos.cpp:499-515,
// data/cube01_hex.mesh:28-53).  The file is reduced to per-axis breakpoints (the input of RectMesh);
// anything else (simplices, curved or non-rectilinear vertices, other attribute layouts) is rejected
// with a message, like the reference aborts on unsupported input.
//
// Both vertex encodings that occur in data/ are read: the plain "vertices / N / dim / coordinates"
// block and the "vertices / N / nodes / FiniteElementSpace ... (Linear | H1_*_P1)" grid function; so is the
// element-wise "L2_T1_<dim>D_P<k>" nodes block that host/mesh_writer.hpp emits (corners -> vertices).
#pragma once
#include <algorithm>
#include <cmath>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace lagb {

inline bool read_mfem_mesh_rectilinear(const std::string &path, int &dim, std::vector<double> coarse[3], std::string &err)
{
   std::ifstream in(path);
   if (!in) { err = "cannot open " + path; return false; }
   std::vector<std::string> tok;
   {
      std::string line;
      bool first = true;
      while (std::getline(in, line))
      {
         if (first) { first = false; if (line.rfind("MFEM mesh v1.0", 0) != 0) { err = "not an MFEM mesh v1.0 file"; return false; } continue; }
         const size_t h = line.find('#');
         if (h != std::string::npos) { line = line.substr(0, h); }
         std::istringstream ls(line);
         std::string t;
         while (ls >> t) { tok.push_back(t); }
      }
   }
   size_t p = 0;
   auto need = [&](const char *kw) -> bool
   {
      while (p < tok.size() && tok[p] != kw) { p++; }
      if (p >= tok.size()) { err = std::string("missing section '") + kw + "'"; return false; }
      p++; return true;
   };
   auto geti = [&](long &v) -> bool { if (p >= tok.size()) { return false; } try { v = std::stol(tok[p++]); } catch (...) { return false; } return true; };
   auto getd = [&](double &v) -> bool { if (p >= tok.size()) { return false; } try { v = std::stod(tok[p++]); } catch (...) { return false; } return true; };
   long d = 0, ne = 0, nb = 0, nv = 0;
   if (!need("dimension") || !geti(d)) { if (err.empty()) { err = "bad dimension"; } return false; }
   if (d != 2 && d != 3) { err = "only 2D / 3D meshes"; return false; }
   dim = (int)d;
   const int nvert_el = 1 << dim, nvert_bd = 1 << (dim - 1);
   const long geom_el = (dim == 2) ? 3 : 5, geom_bd = (dim == 2) ? 1 : 3;   // SQUARE / CUBE, SEGMENT / SQUARE
   if (!need("elements") || !geti(ne)) { if (err.empty()) { err = "bad elements header"; } return false; }
   std::vector<long> ev((size_t)ne*nvert_el);
   for (long e = 0; e < ne; e++)
   {
      long attr, geom;
      if (!geti(attr) || !geti(geom)) { err = "truncated elements"; return false; }
      if (geom != geom_el) { err = "only tensor-product (quad / hex) elements are supported"; return false; }
      for (int k = 0; k < nvert_el; k++) { if (!geti(ev[(size_t)e*nvert_el + k])) { err = "truncated elements"; return false; } }
   }
   if (!need("boundary") || !geti(nb)) { if (err.empty()) { err = "bad boundary header"; } return false; }
   std::vector<long> battr(nb), bv((size_t)nb*nvert_bd);
   for (long b = 0; b < nb; b++)
   {
      long geom;
      if (!geti(battr[b]) || !geti(geom)) { err = "truncated boundary"; return false; }
      if (geom != geom_bd) { err = "unexpected boundary element geometry"; return false; }
      for (int k = 0; k < nvert_bd; k++) { if (!geti(bv[(size_t)b*nvert_bd + k])) { err = "truncated boundary"; return false; } }
   }
   if (!need("vertices") || !geti(nv)) { if (err.empty()) { err = "bad vertices header"; } return false; }
   std::vector<double> X((size_t)nv*dim);
   if (p < tok.size() && tok[p] == "nodes")
   {
      // nodes / FiniteElementSpace / FiniteElementCollection: <name> / VDim: <d> / Ordering: <o>
      std::string fec; long vdim = 0, ordering = 0;
      while (p < tok.size() && tok[p] != "FiniteElementCollection:") { p++; }
      if (p + 1 >= tok.size()) { err = "bad nodes header"; return false; }
      fec = tok[p + 1]; p += 2;
      while (p < tok.size() && tok[p] != "VDim:") { p++; }
      p++; if (!geti(vdim)) { err = "bad VDim"; return false; }
      while (p < tok.size() && tok[p] != "Ordering:") { p++; }
      p++; if (!geti(ordering)) { err = "bad Ordering"; return false; }
      const bool linear = (fec == "Linear") || (fec.find("_P1") != std::string::npos && fec.rfind("H1_", 0) == 0);
      const std::string dg = "L2_T1_" + std::to_string(dim) + "D_P";    // element-wise Gauss-Lobatto (host/mesh_writer.hpp)
      if (fec.rfind(dg, 0) == 0 && vdim == dim)
      {
         // element-major, lexicographic inside an element: the corners give the vertex coordinates; the
         // interior nodes are not checked (a deformed element fails the tensor-grid test below through its corners
         // only if they moved: the reader is for initial / rectilinear meshes)
         long k = 0;
         try { k = std::stol(fec.substr(dg.size())); } catch (...) { k = 0; }
         if (k < 1 || k > 16) { err = "bad nodes collection: " + fec; return false; }
         long nd = 1; for (int a = 0; a < dim; a++) { nd *= k + 1; }
         const long ntot = ne*nd;
         std::vector<double> N((size_t)ntot*dim);
         for (long i = 0; i < ntot*dim; i++)
         {
            double v; if (!getd(v)) { err = "truncated nodes"; return false; }
            const long node = (ordering == 0) ? i % ntot : i / dim, comp = (ordering == 0) ? i / ntot : i % dim;
            N[(size_t)node*dim + comp] = v;
         }
         static const int cq[4][2] = {{0,0},{1,0},{1,1},{0,1}};
         for (long e = 0; e < ne; e++)
            for (int c = 0; c < nvert_el; c++)
            {
               const long vtx = ev[(size_t)e*nvert_el + c];
               if (vtx < 0 || vtx >= nv) { err = "vertex index out of range"; return false; }
               const long loc = k*cq[c % 4][0] + (k + 1)*(k*cq[c % 4][1] + (k + 1)*(k*(long)(c / 4)));
               for (int a = 0; a < dim; a++) { X[(size_t)vtx*dim + a] = N[(size_t)(e*nd + loc)*dim + a]; }
            }
      }
      else
      {
         if (!linear || vdim != dim) { err = "only linear (P1) or element-wise L2_T1 nodal coordinates are supported: " + fec; return false; }
         for (long i = 0; i < nv*dim; i++)
         {
            double v; if (!getd(v)) { err = "truncated nodes"; return false; }
            const long node = (ordering == 0) ? i % nv : i / dim, comp = (ordering == 0) ? i / nv : i % dim;
            X[(size_t)node*dim + comp] = v;
         }
      }
   }
   else
   {
      long vd = 0;
      if (!geti(vd) || vd != dim) { err = "vertex dimension differs from the mesh dimension"; return false; }
      for (long i = 0; i < nv*dim; i++) { if (!getd(X[i])) { err = "truncated vertices"; return false; } }
   }
   // per-axis breakpoints = distinct coordinate values
   double span = 0.0;
   for (int a = 0; a < dim; a++)
   {
      double lo = 1e300, hi = -1e300;
      for (long i = 0; i < nv; i++) { lo = std::min(lo, X[(size_t)i*dim + a]); hi = std::max(hi, X[(size_t)i*dim + a]); }
      span = std::max(span, hi - lo);
   }
   const double tol = 1e-12*std::max(span, 1e-300);
   long ncell = 1;
   for (int a = 0; a < 3; a++) { coarse[a].clear(); }
   for (int a = 0; a < dim; a++)
   {
      std::vector<double> v(nv);
      for (long i = 0; i < nv; i++) { v[i] = X[(size_t)i*dim + a]; }
      std::sort(v.begin(), v.end());
      for (double x : v) { if (coarse[a].empty() || x - coarse[a].back() > tol) { coarse[a].push_back(x); } }
      if (coarse[a].size() < 2) { err = "degenerate mesh"; return false; }
      ncell *= (long)coarse[a].size() - 1;
   }
   if (ncell != ne) { err = "not a rectilinear mesh (element count differs from the product of the axis cells)"; return false; }
   auto index_of = [&](int a, double x) -> int
   {
      const auto it = std::lower_bound(coarse[a].begin(), coarse[a].end(), x - tol);
      if (it == coarse[a].end() || std::fabs(*it - x) > tol) { return -1; }
      return (int)(it - coarse[a].begin());
   };
   // every element must be one cell of the tensor grid (all 2^dim corners, adjacent breakpoints)
   std::vector<char> seen((size_t)ne, 0);
   for (long e = 0; e < ne; e++)
   {
      int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-1, -1, -1};
      for (int k = 0; k < nvert_el; k++)
      {
         const long vtx = ev[(size_t)e*nvert_el + k];
         if (vtx < 0 || vtx >= nv) { err = "vertex index out of range"; return false; }
         for (int a = 0; a < dim; a++)
         {
            const int id = index_of(a, X[(size_t)vtx*dim + a]);
            if (id < 0) { err = "vertex off the tensor grid"; return false; }
            lo[a] = std::min(lo[a], id); hi[a] = std::max(hi[a], id);
         }
      }
      long cell = 0, stride = 1;
      for (int a = 0; a < dim; a++)
      {
         if (hi[a] - lo[a] != 1) { err = "not a rectilinear mesh (element spans more than one cell)"; return false; }
         cell += stride*lo[a]; stride *= (long)coarse[a].size() - 1;
      }
      if (seen[cell]) { err = "not a rectilinear mesh (two elements in one cell)"; return false; }
      seen[cell] = 1;
   }
   // boundary attribute k on faces of constant x_{k-1}, all on the domain boundary
   for (long b = 0; b < nb; b++)
   {
      int axis = -1;
      for (int a = 0; a < dim; a++)
      {
         bool constant = true;
         const double x0 = X[(size_t)bv[(size_t)b*nvert_bd]*dim + a];
         for (int k = 1; k < nvert_bd; k++) { if (std::fabs(X[(size_t)bv[(size_t)b*nvert_bd + k]*dim + a] - x0) > tol) { constant = false; } }
         if (constant && (std::fabs(x0 - coarse[a].front()) <= tol || std::fabs(x0 - coarse[a].back()) <= tol)) { axis = a; }
      }
      if (axis < 0) { err = "boundary element is not on an axis-aligned domain face"; return false; }
      if (battr[b] != axis + 1) { err = "boundary attributes do not follow the convention attribute k = faces of constant x_{k-1}"; return false; }
   }
   return true;
}

} // namespace lagb
