// Writers for the reference's output files: the `-print` files of laghos.cpp:873-900
// (<basename>_<ti>_mesh, _rho, _v, _e) and the VisIt data collection of laghos.cpp:866-871
// (<name>_<cycle>.mfem_root + <name>_<cycle>/{mesh,Density,Velocity,"Specific Internal Energy"}.<rank>),
// in MFEM's text formats (mesh v1.0 and GridFunction::Save).  Host code, off the timed path.
//
// The mesh is written curved: its `nodes` grid function carries the Lagrangian positions x(t).  The
// H1 fields (nodes, velocity) are stored ELEMENT-WISE in the discontinuous Gauss-Lobatto space
// L2_T1_<dim>D_P<ok>: the H1 basis of the path is the Lagrange basis at the same Gauss-Lobatto
// nodes, so the representation is exact, its degree-of-freedom order is fixed by
// the element list alone (element-major, lexicographic inside an element), and the file does not
// depend on a reader's internal edge / face numbering.  The L2 fields (e, rho) are written in the
// path's own space L2_T2_<dim>D_P<ot> (positive / Bernstein basis) and layout, which IS that order.
// Ordering 0 (byNODES): all values of component 0, then component 1, ...
#pragma once
#include "problem.hpp"
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <ostream>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace lagb {

// vertex of element-local corner (cx,cy,cz) in the rank-local vertex lattice
inline long writer_vertex_id(const Problem &P, const int *ie, int cx, int cy, int cz)
{
   const long nvx = P.nloc[0] + 1, nvy = P.nloc[1] + 1;
   return (ie[0] + cx) + nvx*((ie[1] + cy) + nvy*(long)(ie[2] + cz));
}

// MFEM mesh v1.0 of this rank's element block; h_x = H1 position vector (dim x ndofs_h1, byNODES) or null
// for the initial mesh.  Element e of the file is element e of the path (x fastest); boundary elements carry
// attribute k on faces of constant x_{k-1} of the GLOBAL domain (laghos.cpp:499-515), so interior faces of a
// partitioned block are not listed.
inline void write_mfem_mesh(std::ostream &os, const Problem &P, const double *h_x, int precision)
{
   const int dim = P.dim;
   const double *X = h_x ? h_x : P.S0.data();
   os << "MFEM mesh v1.0\n\n"
      << "#\n# MFEM Geometry Types (see mesh/geom.hpp):\n#\n"
      << "# POINT       = 0\n# SEGMENT     = 1\n# TRIANGLE    = 2\n# SQUARE      = 3\n"
      << "# TETRAHEDRON = 4\n# CUBE        = 5\n# PRISM       = 6\n#\n\n";
   os << "dimension\n" << dim << "\n\n";
   // corner order of MFEM's SQUARE / CUBE: counter-clockwise bottom face, then the top face
   static const int cq[4][2] = {{0,0},{1,0},{1,1},{0,1}};
   os << "elements\n" << P.NE << "\n";
   for (int e = 0; e < P.NE; e++)
   {
      const int ie[3] = {e % P.nloc[0], (e/P.nloc[0]) % P.nloc[1], e/(P.nloc[0]*P.nloc[1])};
      os << 1 << ' ' << (dim == 2 ? 3 : 5);
      for (int top = 0; top < (dim == 3 ? 2 : 1); top++)
         for (int k = 0; k < 4; k++) { os << ' ' << writer_vertex_id(P, ie, cq[k][0], cq[k][1], top); }
      os << "\n";
   }
   // boundary faces: outward orientation (counter-clockwise seen from outside)
   std::ostringstream bs;
   long nb = 0;
   for (int a = 0; a < dim; a++)
      for (int side = 0; side < 2; side++)
      {
         if (side == 0 ? (P.lo[a] != 0) : (P.hi[a] != P.mesh.n[a])) { continue; }
         const int b = (a + 1) % dim, c = (a + 2) % dim;      // the face's own axes (3D); b only in 2D
         const int nbb = P.nloc[b], ncc = (dim == 3) ? P.nloc[c] : 1;
         for (int jc = 0; jc < ncc; jc++)
            for (int jb = 0; jb < nbb; jb++)
            {
               int ie[3] = {0, 0, 0};
               ie[a] = (side == 0) ? 0 : P.nloc[a] - 1; ie[b] = jb; if (dim == 3) { ie[c] = jc; }
               auto vid = [&](int ub, int uc) -> long
               {
                  int cc[3] = {0, 0, 0};
                  cc[a] = side; cc[b] = ub; if (dim == 3) { cc[c] = uc; }
                  return writer_vertex_id(P, ie, cc[0], cc[1], cc[2]);
               };
               bs << (a + 1) << ' ' << (dim == 2 ? 1 : 3);
               if (dim == 2)
               {
                  // (a,b) right-handed for a = 0, left-handed for a = 1: keep the domain on the left of the edge
                  const bool fwd = (a == 0) ? (side == 1) : (side == 0);
                  bs << ' ' << vid(fwd ? 0 : 1, 0) << ' ' << vid(fwd ? 1 : 0, 0);
               }
               else
               {
                  // (b,c,a) is a cyclic permutation of (x,y,z): (0,0),(1,0),(1,1),(0,1) in (b,c) has normal +a
                  if (side == 1) { bs << ' ' << vid(0,0) << ' ' << vid(1,0) << ' ' << vid(1,1) << ' ' << vid(0,1); }
                  else           { bs << ' ' << vid(0,0) << ' ' << vid(0,1) << ' ' << vid(1,1) << ' ' << vid(1,0); }
               }
               bs << "\n"; nb++;
            }
      }
   os << "\nboundary\n" << nb << "\n" << bs.str();
   const long nv = (long)(P.nloc[0] + 1)*(P.nloc[1] + 1)*(dim == 3 ? P.nloc[2] + 1 : 1);
   os << "\nvertices\n" << nv << "\n\n";
   os << "nodes\nFiniteElementSpace\nFiniteElementCollection: L2_T1_" << dim << "D_P" << P.spec.ok << "\n"
      << "VDim: " << dim << "\nOrdering: 0\n\n";
   os << std::setprecision(precision);
   for (int c = 0; c < dim; c++)
      for (size_t i = 0; i < (size_t)P.NE*P.ND; i++) { os << X[(size_t)c*P.ndofs_h1 + P.h1_map[i]] << "\n"; }
}

// GridFunction::Save of an H1 field of the path (vdim components, byNODES over the H1 dofs), element-wise
inline void write_h1_field(std::ostream &os, const Problem &P, const double *h_f, int vdim, int precision)
{
   os << "FiniteElementSpace\nFiniteElementCollection: L2_T1_" << P.dim << "D_P" << P.spec.ok << "\n"
      << "VDim: " << vdim << "\nOrdering: 0\n\n";
   os << std::setprecision(precision);
   for (int c = 0; c < vdim; c++)
      for (size_t i = 0; i < (size_t)P.NE*P.ND; i++) { os << h_f[(size_t)c*P.ndofs_h1 + P.h1_map[i]] << "\n"; }
}

// GridFunction::Save of an L2 field of the path (e, rho: [e*NL + l], Bernstein coefficients)
inline void write_l2_field(std::ostream &os, const Problem &P, const double *h_f, int precision)
{
   os << "FiniteElementSpace\nFiniteElementCollection: L2_T2_" << P.dim << "D_P" << P.spec.ot << "\n"
      << "VDim: 1\nOrdering: 0\n\n";
   os << std::setprecision(precision);
   for (int64_t i = 0; i < P.ndofs_l2; i++) { os << h_f[i] << "\n"; }
}

// <collection>_<cycle>.mfem_root of VisItDataCollection::Save (JSON).  As in MFEM's DataCollection the collection name
// may carry a directory prefix ("results/Laghos"): the root file and the per-cycle directory live under the prefix, and
// the paths INSIDE the root file are relative to it ("Laghos_000005/mesh.%06d", 6-digit cycle and rank).  Every field
// is tagged assoc "nodes" with lod = its polynomial order (>= 1); mesh format "0" = serial-format files, one per rank.
// fields: (name, components, order).
struct VisitField { std::string name; int comps, order; };
inline void write_visit_root(std::ostream &os, const std::string &collection, int cycle, double time, double time_step,
                             int nranks, int dim, const std::vector<VisitField> &fields)
{
   char cyc[32]; snprintf(cyc, sizeof(cyc), "%06d", cycle);
   const size_t slash = collection.find_last_of('/');
   const std::string name = (slash == std::string::npos) ? collection : collection.substr(slash + 1);
   const std::string dir = name + "_" + cyc + "/";
   os << "{\n  \"dsets\": {\n    \"main\": {\n"
      << "      \"cycle\": " << cycle << ",\n"
      << "      \"domains\": " << nranks << ",\n"
      << "      \"fields\": {\n";
   for (size_t i = 0; i < fields.size(); i++)
   {
      os << "        \"" << fields[i].name << "\": {\n"
         << "          \"path\": \"" << dir << fields[i].name << ".%06d\",\n"
         << "          \"tags\": { \"assoc\": \"nodes\", \"comps\": \"" << fields[i].comps << "\", \"lod\": \""
         << std::max(1, fields[i].order) << "\" }\n"
         << "        }" << (i + 1 < fields.size() ? "," : "") << "\n";
   }
   os << "      },\n"
      << "      \"mesh\": {\n"
      << "        \"format\": \"0\",\n"
      << "        \"path\": \"" << dir << "mesh.%06d\",\n"
      << "        \"tags\": { \"max_lods\": \"32\", \"spatial_dim\": \"" << dim << "\", \"topo_dim\": \"" << dim << "\" }\n"
      << "      },\n"
      << std::setprecision(16)
      << "      \"time\": " << time << ",\n"
      << "      \"time_step\": " << time_step << "\n"
      << "    }\n  }\n}\n";
}

// ---- files ----
template <class F> inline bool write_text_file(const std::string &path, std::string &err, F &&body)
{
   std::ofstream os(path);
   if (!os) { err = "cannot open " + path + " for writing"; return false; }
   body(os);
   os.flush();
   if (!os) { err = "write failed: " + path; return false; }
   return true;
}

// the four `-print` files of step ti (laghos.cpp:873-900) from the host state S = (x | v | e) and the density rho;
// suffix = "" on one rank, ".<rank:06d>" for the block of one rank of a partitioned run (the reference gathers the
// blocks into one file, PrintAsOne / SaveAsOne; here every rank writes its own block)
inline bool write_print_files(const Problem &P, const std::string &basename, int ti, const double *S, const double *rho,
                              int precision, const std::string &suffix, std::string &err)
{
   const std::string b = basename + "_" + std::to_string(ti);
   const int64_t NV = P.h1_vsize();
   return write_text_file(b + "_mesh" + suffix, err, [&](std::ostream &os) { write_mfem_mesh(os, P, S, precision); })
       && write_text_file(b + "_rho" + suffix, err, [&](std::ostream &os) { write_l2_field(os, P, rho, precision); })
       && write_text_file(b + "_v" + suffix, err, [&](std::ostream &os) { write_h1_field(os, P, S + NV, P.dim, precision); })
       && write_text_file(b + "_e" + suffix, err, [&](std::ostream &os) { write_l2_field(os, P, S + 2*NV, precision); });
}

// one cycle of the VisIt data collection (laghos.cpp:690-698, 866-871): this rank's files and, on rank 0, the root file
inline bool write_visit_files(const Problem &P, const std::string &collection, int cycle, double time, double time_step,
                              int rank, int nranks, const double *S, const double *rho, int precision, std::string &err)
{
   char cyc[32], rk[32];
   snprintf(cyc, sizeof(cyc), "%06d", cycle); snprintf(rk, sizeof(rk), "%06d", rank);
   const std::string dir = collection + "_" + cyc;
   if (mkdir(dir.c_str(), 0777) != 0)
   {
      struct stat st;
      if (stat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) { err = "cannot create directory " + dir; return false; }
   }
   const int64_t NV = P.h1_vsize();
   std::vector<VisitField> fields;
   bool ok = write_text_file(dir + "/mesh." + rk, err, [&](std::ostream &os) { write_mfem_mesh(os, P, S, precision); });
   if (ok && rho)
   {
      fields.push_back({"Density", 1, P.spec.ot});
      ok = write_text_file(dir + "/Density." + rk, err, [&](std::ostream &os) { write_l2_field(os, P, rho, precision); });
   }
   if (ok)
   {
      fields.push_back({"Velocity", P.dim, P.spec.ok});
      ok = write_text_file(dir + "/Velocity." + rk, err, [&](std::ostream &os) { write_h1_field(os, P, S + NV, P.dim, precision); });
   }
   if (ok)
   {
      fields.push_back({"Specific Internal Energy", 1, P.spec.ot});
      ok = write_text_file(dir + "/Specific Internal Energy." + rk, err, [&](std::ostream &os) { write_l2_field(os, P, S + 2*NV, precision); });
   }
   if (ok && rank == 0)
   {
      ok = write_text_file(dir + ".mfem_root", err, [&](std::ostream &os)
      { write_visit_root(os, collection, cycle, time, time_step, nranks, P.dim, fields); });
   }
   return ok;
}

} // namespace lagb
