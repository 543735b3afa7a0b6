// Host-side problem setup: rectilinear hex/quad mesh, H1/L2 numbering, boundary
// (essential) dofs and the eight Laghos initial conditions.
//
// This is setup, not the timed path.  It restates what the reference driver does
// between reading the mesh and constructing LagrangianHydroOperator
// (reference laghos.cpp:380-656): uniform refinement of a Cartesian coarse mesh,
// H1 (Gauss-Lobatto, order ok, vector, Ordering::byNODES) and L2 (Bernstein,
// order ot) spaces, boundary attributes 1/2/3 = faces of constant x/y/z
// (data/cube01_hex.mesh:28-53, laghos.cpp:499-515), nodal velocity projection
// (laghos.cpp:574-579), nodal-L2 -> Bernstein projection of rho0 and e
// (laghos.cpp:589-595, 617-622), the Sedov delta function (laghos.cpp:597-606) and
// the piecewise-constant gamma (laghos.cpp:628-632).  All gate meshes of the
// reference (square01_quad, cube01_hex, box01_hex, rectangle01_quad,
// square_gresho, rt2D) are rectilinear, so the mesh is generated from per-axis
// breakpoints instead of being read from a file.
//
// Element numbering is lexicographic (x fastest); MFEM's refinement numbering is
// different, which only changes floating-point summation order.
#pragma once
#include "fe_tables.hpp"
#include <cstdint>
#include <cstdio>
#include <string>

namespace lagb {

struct RectMesh
{
   int dim = 3;
   int n[3] = {1, 1, 1};
   std::vector<double> brk[3];

   // coarse: breakpoints per axis; rs: number of uniform refinements.
   void build(int dim_, const std::vector<double> coarse[3], int rs)
   {
      dim = dim_;
      for (int d = 0; d < 3; d++)
      {
         brk[d] = (d < dim) ? coarse[d] : std::vector<double> {0.0, 1.0};
         if (d < dim)
         {
            for (int l = 0; l < rs; l++)
            {
               std::vector<double> r;
               for (size_t i = 0; i + 1 < brk[d].size(); i++)
               {
                  r.push_back(brk[d][i]);
                  r.push_back(0.5*(brk[d][i] + brk[d][i+1]));
               }
               r.push_back(brk[d].back());
               brk[d].swap(r);
            }
         }
         n[d] = (d < dim) ? (int)brk[d].size() - 1 : 1;
      }
   }
   int NE() const { return n[0]*n[1]*n[2]; }
};

// Named coarse meshes of the reference's data/ directory.
inline bool named_coarse_mesh(const std::string &name, int &dim,
                              std::vector<double> coarse[3])
{
   coarse[0].clear(); coarse[1].clear(); coarse[2].clear();
   if (name == "square01_quad")
   { dim = 2; coarse[0] = {0, .5, 1}; coarse[1] = {0, .5, 1}; return true; }
   if (name == "cube01_hex")
   {
      dim = 3; coarse[0] = {0, .5, 1}; coarse[1] = {0, .5, 1}; coarse[2] = {0, .5, 1};
      return true;
   }
   if (name == "box01_hex")
   {
      dim = 3; coarse[0] = {0, 1, 3, 5, 7}; coarse[1] = {0, 1.5, 3};
      coarse[2] = {0, 1.5, 3}; return true;
   }
   if (name == "rectangle01_quad")
   {
      dim = 2; coarse[0] = {0, 1, 2, 3, 4, 5, 6, 7}; coarse[1] = {0, 1, 2, 3};
      return true;
   }
   if (name == "square_gresho")
   { dim = 2; coarse[0] = {-.5, 0, .5}; coarse[1] = {-.5, 0, .5}; return true; }
   if (name == "rt2D")
   { dim = 2; coarse[0] = {0, .5}; coarse[1] = {-1, -.5, 0, .5, 1}; return true; }
   // weak-scaling family (not a reference mesh): hexbox_PxQxR = P x Q x R coarse hexes of
   // edge 0.5 starting at the origin, so hexbox_2x2x2 == cube01_hex and every member has
   // the same element size at equal -rs (bench.py --gpus 2 / 4).
   int p = 0, q = 0, r = 0;
   if (sscanf(name.c_str(), "hexbox_%dx%dx%d", &p, &q, &r) == 3 && p > 0 && q > 0 && r > 0 && p <= 64 && q <= 64 && r <= 64)
   {
      dim = 3;
      const int n[3] = {p, q, r};
      for (int d = 0; d < 3; d++) { for (int i = 0; i <= n[d]; i++) { coarse[d].push_back(0.5*i); } }
      return true;
   }
   // the reference's built-in mesh (`-m default`, laghos.cpp:131-137, 427-447: Mesh::MakeCartesian2D/3D(nx, ny, nz, Sx, Sy,
   // Sz) with boundary attribute k on the faces of constant x_{k-1}): "default" = -dim 3 -nx 2 -ny 2 -nz 2 on the unit
   // cube (the mesh of the --checks table, = cube01_hex), "default_2d" = -dim 2 (= square01_quad), and
   // "default_<nx>x<ny>[x<nz>][_S<Sx>x<Sy>[x<Sz>]]" for -nx/-ny/-nz and -Sx/-Sy/-Sz
   if (name.rfind("default", 0) == 0)
   {
      int n[3] = {2, 2, 2}; double S[3] = {1.0, 1.0, 1.0};
      dim = 3;
      const std::string rest = name.substr(7);
      if (rest == "_2d") { dim = 2; }
      else if (!rest.empty())
      {
         char tail[64] = "";
         int a = 0, b = 0, c = 0;
         const int got = sscanf(rest.c_str(), "_%dx%dx%d%63s", &a, &b, &c, tail);
         if (got >= 3) { dim = 3; n[0] = a; n[1] = b; n[2] = c; }
         else if (sscanf(rest.c_str(), "_%dx%d%63s", &a, &b, tail) >= 2) { dim = 2; n[0] = a; n[1] = b; }
         else { return false; }
         if (tail[0] != 0)
         {
            double sx = 0, sy = 0, sz = 0; char junk = 0;
            if (dim == 3) { if (sscanf(tail, "_S%lfx%lfx%lf%c", &sx, &sy, &sz, &junk) != 3) { return false; } }
            else { sz = 1.0; if (sscanf(tail, "_S%lfx%lf%c", &sx, &sy, &junk) != 2) { return false; } }
            if (!(sx > 0 && sy > 0 && sz > 0)) { return false; }
            S[0] = sx; S[1] = sy; S[2] = sz;
         }
      }
      for (int d = 0; d < dim; d++)
      {
         if (n[d] < 1 || n[d] > 4096) { return false; }
         for (int i = 0; i <= n[d]; i++) { coarse[d].push_back(i == n[d] ? S[d] : S[d]*i/n[d]); }
      }
      return true;
   }
   return false;
}

struct ProblemSpec
{
   int problem = 1;        // reference -p
   int dim = 3;
   int ok = 2, ot = 1, oq = -1;
   double blast_scale = 0.125; // value given to DeltaCoefficient: E0/2^dim
   // (parallel driver, laghos.cpp:603-604) or 0.25 (serial/laghos.cpp:101,319)
   bool impose_visc = false;
};

// Initial-condition functions: reference laghos.cpp:1094-1275.
struct ICs
{
   int problem, dim;
   double rho0(const double *x) const
   {
      switch (problem)
      {
         case 0: return 1.0;
         case 1: return 1.0;
         case 2: return (x[0] < 0.5) ? 1.0 : 0.1;
         case 3: return (dim == 2) ? (x[0] > 1.0 && x[1] > 1.5) ? 0.125 : 1.0
                           : x[0] > 1.0 && ((x[1] < 1.5 && x[2] < 1.5) ||
                                            (x[1] > 1.5 && x[2] > 1.5)) ? 0.125 : 1.0;
         case 4: return 1.0;
         case 5:
            if (x[0] >= 0.5 && x[1] >= 0.5) { return 0.5313; }
            if (x[0] <  0.5 && x[1] <  0.5) { return 0.8; }
            return 1.0;
         case 6:
            if (x[0] <  0.5 && x[1] >= 0.5) { return 2.0; }
            if (x[0] >= 0.5 && x[1] <  0.5) { return 3.0; }
            return 1.0;
         case 7: return x[1] >= 0.0 ? 2.0 : 1.0;
      }
      return 0.0;
   }
   double gamma(const double *x) const
   {
      switch (problem)
      {
         case 0: return 5.0/3.0;
         case 1: return 1.4;
         case 2: return 1.4;
         case 3: return (x[0] > 1.0 && x[1] <= 1.5) ? 1.4 : 1.5;
         case 4: return 5.0/3.0;
         case 5: return 1.4;
         case 6: return 1.4;
         case 7: return 5.0/3.0;
      }
      return 0.0;
   }
   void v0(const double *x, double *v) const
   {
      const double atn = pow((x[0]*(1.0 - x[0])*4*x[1]*(1.0 - x[1])*4.0), 0.4);
      for (int d = 0; d < dim; d++) { v[d] = 0.0; }
      switch (problem)
      {
         case 0:
            v[0] =  sin(M_PI*x[0])*cos(M_PI*x[1]);
            v[1] = -cos(M_PI*x[0])*sin(M_PI*x[1]);
            if (dim == 3)
            {
               v[0] *= cos(M_PI*x[2]);
               v[1] *= cos(M_PI*x[2]);
               v[2] = 0.0;
            }
            break;
         case 1: case 2: case 3: break;
         case 4:
         {
            const double r = sqrt(x[0]*x[0] + x[1]*x[1]);
            if (r < 0.2)
            {
               v[0] =  5.0*x[1];
               v[1] = -5.0*x[0];
            }
            else if (r < 0.4)
            {
               v[0] =  2.0*x[1]/r - 5.0*x[1];
               v[1] = -2.0*x[0]/r + 5.0*x[0];
            }
            break;
         }
         case 5:
            if (x[0] >= 0.5 && x[1] >= 0.5) { v[0] = 0.0*atn; v[1] = 0.0*atn; return; }
            if (x[0] <  0.5 && x[1] >= 0.5) { v[0] = 0.7276*atn; v[1] = 0.0*atn; return; }
            if (x[0] <  0.5 && x[1] <  0.5) { v[0] = 0.0*atn; v[1] = 0.0*atn; return; }
            if (x[0] >= 0.5 && x[1] <  0.5) { v[0] = 0.0*atn; v[1] = 0.7276*atn; return; }
            break;
         case 6:
            if (x[0] >= 0.5 && x[1] >= 0.5) { v[0] = +0.75*atn; v[1] = -0.5*atn; return; }
            if (x[0] <  0.5 && x[1] >= 0.5) { v[0] = +0.75*atn; v[1] = +0.5*atn; return; }
            if (x[0] <  0.5 && x[1] <  0.5) { v[0] = -0.75*atn; v[1] = +0.5*atn; return; }
            if (x[0] >= 0.5 && x[1] <  0.5) { v[0] = -0.75*atn; v[1] = -0.5*atn; return; }
            break;
         case 7:
            v[1] = 0.02*exp(-2*M_PI*x[1]*x[1])*cos(2*M_PI*x[0]);
            break;
      }
   }
   double e0(const double *x) const
   {
      switch (problem)
      {
         case 0:
         {
            const double denom = 2.0/3.0;
            double val;
            if (dim == 2)
            {
               val = 1.0 + (cos(2*M_PI*x[0]) + cos(2*M_PI*x[1]))/4.0;
            }
            else
            {
               val = 100.0 + ((cos(2*M_PI*x[2]) + 2)*
                              (cos(2*M_PI*x[0]) + cos(2*M_PI*x[1])) - 2)/16.0;
            }
            return val/denom;
         }
         case 1: return 0.0;
         case 2: return (x[0] < 0.5) ? 1.0/rho0(x)/(gamma(x) - 1.0)
                           : 0.1/rho0(x)/(gamma(x) - 1.0);
         case 3: return (x[0] > 1.0) ? 0.1/rho0(x)/(gamma(x) - 1.0)
                           : 1.0/rho0(x)/(gamma(x) - 1.0);
         case 4:
         {
            const double r = sqrt(x[0]*x[0] + x[1]*x[1]), rsq = x[0]*x[0] + x[1]*x[1];
            const double gam = 5.0/3.0;
            if (r < 0.2)
            {
               return (5.0 + 25.0/2.0*rsq)/(gam - 1.0);
            }
            else if (r < 0.4)
            {
               const double t1 = 9.0 - 4.0*log(0.2) + 25.0/2.0*rsq;
               const double t2 = 20.0*r - 4.0*log(r);
               return (t1 - t2)/(gam - 1.0);
            }
            else { return (3.0 + 4.0*log(2.0))/(gam - 1.0); }
         }
         case 5:
         {
            const double irg = 1.0/rho0(x)/(gamma(x) - 1.0);
            if (x[0] >= 0.5 && x[1] >= 0.5) { return 0.4*irg; }
            return 1.0*irg;
         }
         case 6:
         {
            const double irg = 1.0/rho0(x)/(gamma(x) - 1.0);
            return 1.0*irg;
         }
         case 7:
         {
            const double rho = rho0(x), gam = gamma(x);
            return (6.0 - rho*x[1])/(gam - 1.0)/rho;
         }
      }
      return 0.0;
   }
};

// Everything the operators need, in the layouts of SURVEY.md section 8(a).
struct Problem
{
   ProblemSpec spec;
   RectMesh mesh;          // the GLOBAL mesh
   int lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};   // this rank's element box [lo,hi) per axis
   int nloc[3] = {1, 1, 1};                    // hi - lo
   Tables1D tab;
   int dim = 3, NE = 0, D1D = 0, L1D = 0, Q1D = 0;
   int ND = 0;      // H1 dofs per element  D1D^dim
   int NL = 0;      // L2 dofs per element  L1D^dim
   int NQ = 0;      // quad points per element
   int N1[3] = {1, 1, 1};  // H1 lattice extents
   int64_t ndofs_h1 = 0;   // scalar H1 dofs
   int64_t ndofs_l2 = 0;
   std::vector<int> h1_map;          // [e*ND + i], lexicographic local i
   std::vector<int> ess[3];          // scalar dof ids with v_c = 0
   std::vector<double> qweights;     // tensor weights [NQ]
   std::vector<double> S0;           // (x | v | e)
   std::vector<double> rho0_gf;      // Bernstein coefficients of rho0, L2 layout
   std::vector<double> rho0_q;       // analytic rho0 at quad points [e*NQ+q] (mass coefficient)
   std::vector<double> gamma;        // per element
   bool use_visc = true, use_vort = false;
   int source = 0;

   int64_t h1_vsize() const { return dim*ndofs_h1; }
   int64_t s_size() const { return 2*h1_vsize() + ndofs_l2; }

   static int ipow(int a, int b) { int r = 1; while (b-- > 0) { r *= a; } return r; }

   // global element index (per axis) of local element e
   void elem_idx(int e, int *idx) const
   {
      idx[0] = lo[0] + e % nloc[0]; idx[1] = lo[1] + (e/nloc[0]) % nloc[1];
      idx[2] = lo[2] + e/(nloc[0]*nloc[1]);
   }
   // physical box of local element e
   void elem_box(int e, double *blo, double *bhi) const
   {
      int idx[3]; elem_idx(e, idx);
      for (int d = 0; d < 3; d++)
      {
         blo[d] = mesh.brk[d][idx[d]]; bhi[d] = mesh.brk[d][idx[d]+1];
      }
   }

   // whole mesh on one rank
   void build(const ProblemSpec &sp, const RectMesh &m)
   {
      const int l0[3] = {0, 0, 0};
      build(sp, m, l0, m.n);
   }

   // element box [lo_,hi_) of the global mesh m (element-partitioned ranks):
   // boundary conditions and the Sedov delta refer to the GLOBAL mesh.
   void build(const ProblemSpec &sp, const RectMesh &m, const int *lo_, const int *hi_)
   {
      spec = sp; mesh = m; dim = m.dim;
      for (int d = 0; d < 3; d++)
      {
         lo[d] = (d < dim) ? lo_[d] : 0; hi[d] = (d < dim) ? hi_[d] : 1; nloc[d] = hi[d] - lo[d];
         if (nloc[d] < 1 || lo[d] < 0 || hi[d] > m.n[d]) { throw std::runtime_error("bad element box"); }
      }
      tab.build(sp.ok, sp.ot, sp.oq);
      D1D = tab.D1D; L1D = tab.L1D; Q1D = tab.Q1D;
      ND = ipow(D1D, dim); NL = ipow(L1D, dim); NQ = ipow(Q1D, dim);
      NE = nloc[0]*nloc[1]*nloc[2];
      for (int d = 0; d < 3; d++) { N1[d] = (d < dim) ? nloc[d]*sp.ok + 1 : 1; }
      ndofs_h1 = (int64_t)N1[0]*N1[1]*N1[2];
      ndofs_l2 = (int64_t)NE*NL;
      if (ndofs_h1*dim > 2000000000LL) { throw std::runtime_error("mesh too large for int32 dof ids"); }

      // physics switches: reference laghos.cpp:635-648
      switch (sp.problem)
      {
         case 0: if (dim == 2) { source = 1; } use_visc = false; break;
         case 4: use_visc = false; break;
         case 7: source = 2; use_visc = true; use_vort = true; break;
         default: use_visc = true;
      }
      if (sp.impose_visc) { use_visc = true; }

      // gather map (MFEM ElementRestriction, LEXICOGRAPHIC ordering)
      h1_map.resize((size_t)NE*ND);
      const int DZ = (dim == 3) ? D1D : 1;
      for (int e = 0; e < NE; e++)
      {
         const int ix = e % nloc[0], iy = (e/nloc[0]) % nloc[1];
         const int iz = e/(nloc[0]*nloc[1]);
         for (int kz = 0; kz < DZ; kz++)
            for (int ky = 0; ky < D1D; ky++)
               for (int kx = 0; kx < D1D; kx++)
               {
                  const int gx = ix*sp.ok + kx, gy = iy*sp.ok + ky, gz = iz*sp.ok + kz;
                  h1_map[(size_t)e*ND + kx + D1D*(ky + D1D*kz)] = gx + N1[0]*(gy + N1[1]*gz);
               }
      }
      // essential dofs: component c vanishes on faces of constant x_c
      for (int c = 0; c < dim; c++)
      {
         ess[c].clear();
         for (int gz = 0; gz < N1[2]; gz++)
            for (int gy = 0; gy < N1[1]; gy++)
               for (int gx = 0; gx < N1[0]; gx++)
               {
                  const int g[3] = {gx, gy, gz};
                  const int gg = lo[c]*sp.ok + g[c];   // global lattice index
                  if (gg == 0 || gg == mesh.n[c]*sp.ok) { ess[c].push_back(gx + N1[0]*(gy + N1[1]*gz)); }
               }
      }
      // tensor quadrature weights, q = qx + Q1D*(qy + Q1D*qz)
      qweights.assign(NQ, 0.0);
      for (int q = 0; q < NQ; q++)
      {
         int r = q; double w = 1.0;
         for (int d = 0; d < dim; d++) { w *= tab.qw[r % Q1D]; r /= Q1D; }
         qweights[q] = w;
      }

      ICs ic {sp.problem, dim};
      S0.assign((size_t)s_size(), 0.0);
      double *X = S0.data(), *V = S0.data() + h1_vsize(), *E = S0.data() + 2*h1_vsize();
      // mesh nodes and nodal velocity
      for (int gz = 0; gz < N1[2]; gz++)
         for (int gy = 0; gy < N1[1]; gy++)
            for (int gx = 0; gx < N1[0]; gx++)
            {
               const int g[3] = {gx, gy, gz};
               double x[3] = {0, 0, 0}, v[3] = {0, 0, 0};
               bool on_bdr[3] = {false, false, false};
               for (int d = 0; d < dim; d++)
               {
                  const int ell = std::min(g[d]/sp.ok, nloc[d] - 1);
                  const int j = g[d] - ell*sp.ok;
                  const int el = lo[d] + ell;
                  const double a = mesh.brk[d][el], b = mesh.brk[d][el+1];
                  x[d] = (j == 0) ? a : (j == sp.ok) ? b : a + (b - a)*tab.gll[j];
                  const int gg = lo[d]*sp.ok + g[d];
                  on_bdr[d] = (gg == 0 || gg == mesh.n[d]*sp.ok);
               }
               ic.v0(x, v);
               const int64_t id = gx + (int64_t)N1[0]*(gy + (int64_t)N1[1]*gz);
               for (int d = 0; d < dim; d++)
               {
                  X[d*ndofs_h1 + id] = x[d];
                  V[d*ndofs_h1 + id] = on_bdr[d] ? 0.0 : v[d];
               }
            }

      // L2 fields: nodal (Gauss-Legendre) values, then change of basis to Bernstein
      rho0_gf.assign((size_t)ndofs_l2, 0.0);
      rho0_q.assign((size_t)NE*NQ, 0.0);
      gamma.assign(NE, 0.0);
      std::vector<double> nod_rho(NL), nod_e(NL);
      const int LZ = (dim == 3) ? L1D : 1, QZ = (dim == 3) ? Q1D : 1;

      // Sedov: vertex closest to the origin and the elements owning it
      // (MFEM GridFunction::ProjectDeltaCoefficient, not in tree)
      int vnear[3] = {0, 0, 0};
      double dist2 = 0.0;
      for (int d = 0; d < dim; d++)
      {
         double best = 1e300;
         for (size_t i = 0; i < mesh.brk[d].size(); i++)
         {
            if (fabs(mesh.brk[d][i]) < best) { best = fabs(mesh.brk[d][i]); vnear[d] = (int)i; }
         }
         dist2 += best*best;
      }
      // the reference only accepts a vertex within -dtol (default 1e-12, laghos.cpp:147, :605) of the blast position
      if (sp.problem == 1 && !(sqrt(dist2) <= 1e-12))
      { throw std::runtime_error("Delta function could not be initialized: no mesh vertex within delta_tol of the blast position"); }
      // unit-weight mass integral of the nodal interpolant over ALL (global) elements
      // that own the vertex: sum_q w detJ f(q); f is a tensor polynomial of degree ot,
      // so Gauss-Legendre(Q1D) is exact.
      double delta_integral = 0.0;
      if (sp.problem == 1)
      {
         const int ncand = 1 << dim;
         for (int k = 0; k < ncand; k++)
         {
            double integ = 1.0; bool ok_el = true;
            for (int d = 0; d < dim; d++)
            {
               const int side = (k >> d) & 1;            // 0: vertex at xi=0, 1: at xi=1
               const int el = vnear[d] - side;
               if (el < 0 || el >= mesh.n[d]) { ok_el = false; break; }
               double s1 = 0.0;
               for (int q = 0; q < Q1D; q++)
               {
                  const double xi = tab.qx[q];
                  s1 += tab.qw[q]*pow(side ? xi : 1.0 - xi, (double)sp.ot);
               }
               integ *= (mesh.brk[d][el+1] - mesh.brk[d][el])*s1;
            }
            if (ok_el) { delta_integral += integ; }
         }
      }
      std::vector<int> delta_side((size_t)NE*3, 0); // 0: vertex at xi=0, 1: at xi=1

      for (int e = 0; e < NE; e++)
      {
         double lo[3], hi[3]; elem_box(e, lo, hi);
         int idx[3]; elem_idx(e, idx);
         double xc[3] = {0, 0, 0};
         for (int d = 0; d < dim; d++) { xc[d] = lo[d] + (hi[d] - lo[d])*0.5; }
         gamma[e] = ic.gamma(xc);

         bool has_vertex = (sp.problem == 1);
         for (int d = 0; d < dim && has_vertex; d++)
         {
            if (idx[d] == vnear[d]) { delta_side[(size_t)e*3 + d] = 0; }
            else if (idx[d] + 1 == vnear[d]) { delta_side[(size_t)e*3 + d] = 1; }
            else { has_vertex = false; }
         }

         for (int lz = 0; lz < LZ; lz++)
            for (int ly = 0; ly < L1D; ly++)
               for (int lx = 0; lx < L1D; lx++)
               {
                  const int l[3] = {lx, ly, lz};
                  double x[3] = {0, 0, 0};
                  for (int d = 0; d < dim; d++) { x[d] = lo[d] + (hi[d] - lo[d])*tab.gl_l2[l[d]]; }
                  const int li = lx + L1D*(ly + L1D*lz);
                  nod_rho[li] = ic.rho0(x);
                  if (sp.problem == 1)
                  {
                     double val = 0.0;
                     if (has_vertex)
                     {
                        val = 1.0;
                        for (int d = 0; d < dim; d++)
                        {
                           const double xi = tab.gl_l2[l[d]];
                           val *= pow(delta_side[(size_t)e*3 + d] ? xi : 1.0 - xi, (double)sp.ot);
                        }
                     }
                     nod_e[li] = val;
                  }
                  else { nod_e[li] = ic.e0(x); }
               }
         nodal_to_bernstein(nod_rho.data(), rho0_gf.data() + (size_t)e*NL);
         nodal_to_bernstein(nod_e.data(), E + (size_t)e*NL);

         for (int qz = 0; qz < QZ; qz++)
            for (int qy = 0; qy < Q1D; qy++)
               for (int qx = 0; qx < Q1D; qx++)
               {
                  const int qq[3] = {qx, qy, qz};
                  double x[3] = {0, 0, 0};
                  for (int d = 0; d < dim; d++) { x[d] = lo[d] + (hi[d] - lo[d])*tab.qx[qq[d]]; }
                  rho0_q[(size_t)e*NQ + qx + Q1D*(qy + Q1D*qz)] = ic.rho0(x);
               }
      }
      if (sp.problem == 1)
      {
         if (!(delta_integral > 0.0)) { throw std::runtime_error("Delta function could not be initialized!"); }
         const double s = sp.blast_scale/delta_integral;
         for (int64_t i = 0; i < ndofs_l2; i++) { E[i] *= s; }
      }
   }

   // Tensor change of basis nodal(GL) -> Bernstein on one element.
   void nodal_to_bernstein(const double *nod, double *bern) const
   {
      const int LZ = (dim == 3) ? L1D : 1;
      std::vector<double> a(nod, nod + NL), b(NL);
      for (int axis = 0; axis < dim; axis++)
      {
         for (int k = 0; k < LZ; k++)
            for (int j = 0; j < L1D; j++)
               for (int i = 0; i < L1D; i++)
               {
                  const int idx[3] = {i, j, k};
                  double s = 0.0;
                  for (int m = 0; m < L1D; m++)
                  {
                     int src[3] = {i, j, k}; src[axis] = m;
                     s += tab.N2B[idx[axis] + L1D*m]*a[src[0] + L1D*(src[1] + L1D*src[2])];
                  }
                  b[i + L1D*(j + L1D*k)] = s;
               }
         a.swap(b);
      }
      for (int i = 0; i < NL; i++) { bern[i] = a[i]; }
   }
};

} // namespace lagb
