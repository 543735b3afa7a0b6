// C ABI for the host-side problem setup (include/laghos_b200.h, "Host-side problem setup").
#include "../../include/laghos_b200.h"
#include "host/problem.hpp"
#include "host/partition.hpp"
#include "host/mesh_reader.hpp"
#include "host/mesh_writer.hpp"
#include "host/error_norms.hpp"
#include "host/checks.hpp"
#include <string>

namespace lagb { void set_error(const std::string &msg); }

struct lagb_problem { lagb::Problem P; lagb::Partition part; bool partitioned = false; };

extern "C" {

int lagb_problem_create_rect(lagb_problem **out, int dim,
                             const double *bx, int nbx, const double *by, int nby,
                             const double *bz, int nbz, int rs, int problem,
                             int ok, int ot, int oq, double blast_scale, int impose_visc)
{
   try
   {
      if (dim != 2 && dim != 3) { lagb::set_error("problem_create: dim must be 2 or 3 (PA path, reference laghos.cpp:454-462)"); return LAGB_ERR_INVALID; }
      if (problem < 0 || problem > 7) { lagb::set_error("Wrong problem specification!"); return LAGB_ERR_INVALID; }
      std::vector<double> coarse[3];
      coarse[0].assign(bx, bx + nbx);
      coarse[1].assign(by, by + nby);
      if (dim == 3) { coarse[2].assign(bz, bz + nbz); }
      lagb::RectMesh rm; rm.build(dim, coarse, rs);
      lagb::ProblemSpec sp;
      sp.problem = problem; sp.dim = dim; sp.ok = ok; sp.ot = ot; sp.oq = oq;
      sp.blast_scale = blast_scale; sp.impose_visc = impose_visc != 0;
      lagb_problem *p = new lagb_problem();
      p->P.build(sp, rm);
      *out = p;
      return LAGB_OK;
   }
   catch (const std::exception &e) { lagb::set_error(e.what()); return LAGB_ERR_INVALID; }
}

int lagb_problem_create(lagb_problem **out, const char *mesh_name, int rs, int problem,
                        int ok, int ot, int oq, double blast_scale, int impose_visc)
{
   std::vector<double> coarse[3]; int dim = 0;
   if (!mesh_name || !lagb::named_coarse_mesh(mesh_name, dim, coarse))
   { lagb::set_error(std::string("unknown mesh: ") + (mesh_name ? mesh_name : "(null)")); return LAGB_ERR_INVALID; }
   return lagb_problem_create_rect(out, dim, coarse[0].data(), (int)coarse[0].size(),
                                   coarse[1].data(), (int)coarse[1].size(),
                                   coarse[2].data(), (int)coarse[2].size(), rs, problem, ok, ot, oq,
                                   blast_scale, impose_visc);
}

int lagb_problem_create_file(lagb_problem **out, const char *path, int rs, int problem,
                             int ok, int ot, int oq, double blast_scale, int impose_visc)
{
   std::vector<double> coarse[3]; int dim = 0; std::string err;
   if (!path || !lagb::read_mfem_mesh_rectilinear(path, dim, coarse, err))
   { lagb::set_error(std::string("mesh file: ") + (path ? err : "(null path)")); return LAGB_ERR_INVALID; }
   return lagb_problem_create_rect(out, dim, coarse[0].data(), (int)coarse[0].size(),
                                   coarse[1].data(), (int)coarse[1].size(),
                                   coarse[2].data(), (int)coarse[2].size(), rs, problem, ok, ot, oq,
                                   blast_scale, impose_visc);
}

int lagb_problem_mesh_breaks(const lagb_problem *p, int axis, const double **brk, int32_t *n)
{
   if (!p || axis < 0 || axis > 2) { return LAGB_ERR_INVALID; }
   *brk = p->P.mesh.brk[axis].data(); *n = (int32_t)p->P.mesh.brk[axis].size();
   return LAGB_OK;
}

int lagb_problem_create_part(lagb_problem **out, const char *mesh_name, int rs, int problem,
                             int ok, int ot, int oq, double blast_scale, int impose_visc,
                             int rank, const int32_t pgrid[3])
{
   try
   {
      std::vector<double> coarse[3]; int dim = 0;
      if (!mesh_name || !lagb::named_coarse_mesh(mesh_name, dim, coarse))
      { lagb::set_error(std::string("unknown mesh: ") + (mesh_name ? mesh_name : "(null)")); return LAGB_ERR_INVALID; }
      lagb::RectMesh gm; gm.build(dim, coarse, rs);
      lagb::ProblemSpec sp;
      sp.problem = problem; sp.dim = dim; sp.ok = ok; sp.ot = ot; sp.oq = oq;
      sp.blast_scale = blast_scale; sp.impose_visc = impose_visc != 0;
      lagb_problem *p = new lagb_problem();
      const int pg[3] = {pgrid[0], pgrid[1], pgrid[2]};
      p->part.build(dim, gm.n, pg, rank, ok);
      p->P.build(sp, gm, p->part.lo, p->part.hi);
      p->partitioned = true;
      *out = p;
      return LAGB_OK;
   }
   catch (const std::exception &e) { lagb::set_error(e.what()); return LAGB_ERR_INVALID; }
}
int lagb_problem_nnbr(const lagb_problem *p) { return p->partitioned ? (int)p->part.nbrs.size() : 0; }
int lagb_problem_nbr(const lagb_problem *p, int k, int32_t *rank, int32_t *phase, int32_t *n, const int32_t **dofs)
{
   if (!p->partitioned || k < 0 || k >= (int)p->part.nbrs.size()) { lagb::set_error("problem_nbr: bad index"); return LAGB_ERR_INVALID; }
   const auto &nb = p->part.nbrs[k];
   *rank = nb.rank; *phase = nb.phase; *n = (int32_t)nb.dofs.size(); *dofs = nb.dofs.data();
   return LAGB_OK;
}
const uint8_t *lagb_problem_owner_mask(const lagb_problem *p) { return p->partitioned ? p->part.owner.data() : nullptr; }

void lagb_problem_destroy(lagb_problem *p) { delete p; }

int lagb_problem_get_info(const lagb_problem *p, lagb_problem_info *o)
{
   const lagb::Problem &P = p->P;
   o->dim = P.dim; o->NE = P.NE; o->D1D = P.D1D; o->L1D = P.L1D; o->Q1D = P.Q1D;
   o->ND = P.ND; o->NL = P.NL; o->NQ = P.NQ;
   for (int d = 0; d < 3; d++) { o->nelem[d] = P.mesh.n[d]; o->n1[d] = P.N1[d]; o->ness[d] = (d < P.dim) ? (int)P.ess[d].size() : 0; }
   o->ndofs_h1 = P.ndofs_h1; o->ndofs_l2 = P.ndofs_l2;
   o->use_visc = P.use_visc; o->use_vort = P.use_vort; o->source = P.source;
   return LAGB_OK;
}
const int32_t *lagb_problem_h1_map(const lagb_problem *p) { return p->P.h1_map.data(); }
const int32_t *lagb_problem_ess(const lagb_problem *p, int c) { return (c >= 0 && c < p->P.dim) ? p->P.ess[c].data() : nullptr; }
const double *lagb_problem_S0(const lagb_problem *p) { return p->P.S0.data(); }
const double *lagb_problem_rho0_gf(const lagb_problem *p) { return p->P.rho0_gf.data(); }
const double *lagb_problem_rho0_q(const lagb_problem *p) { return p->P.rho0_q.data(); }
const double *lagb_problem_gamma(const lagb_problem *p) { return p->P.gamma.data(); }
const double *lagb_problem_qweights(const lagb_problem *p) { return p->P.qweights.data(); }
const double *lagb_problem_table(const lagb_problem *p, int which)
{
   switch (which)
   {
      case 0: return p->P.tab.B.data(); case 1: return p->P.tab.G.data(); case 2: return p->P.tab.BL.data();
      case 3: return p->P.tab.qx.data(); case 4: return p->P.tab.qw.data();
   }
   return nullptr;
}

int lagb_checks_entry(int dim, int problem, int k, int32_t *it, double *norm)
{
   if (dim < 2 || dim > 3 || problem < 0 || problem > 7 || k < 0 || k > 1 || !it || !norm)
   { lagb::set_error("checks_entry: bad argument"); return LAGB_ERR_INVALID; }
   const lagb::CheckEntry &c = lagb::checks_table[(dim - 2)*8 + problem][k];
   *it = c.it; *norm = c.norm;
   return LAGB_OK;
}

int lagb_checks_step(int dim, int problem, int ti, double e_norm, double eps, int32_t *chk)
{
   if (!chk) { lagb::set_error("checks_step: null argument"); return LAGB_ERR_INVALID; }
   std::string msg; int n = *chk;
   const int rc = lagb::checks_step(dim, problem, ti, e_norm, eps, n, msg);
   *chk = n;
   if (rc < 0) { lagb::set_error(msg); return LAGB_ERR_INVALID; }
   return LAGB_OK;
}

int lagb_problem_velocity_error(const lagb_problem *p, const double *h_S, double out[4])
{
   if (!p || !h_S || !out) { lagb::set_error("velocity_error: null argument"); return LAGB_ERR_INVALID; }
   double s[3];
   lagb::velocity_error_sums(p->P, h_S, s);
   out[0] = s[0]; out[1] = s[1]; out[2] = std::sqrt(s[2]); out[3] = s[2];
   return LAGB_OK;
}

int lagb_sedov_exact_eval(int dim, double gamma, double rho0, double blast_energy, double t, double alpha_override,
                          int n, const double *r, double *rho, double *v, double *pr, double info[6])
{
   try
   {
      if (!(t > 0) || n < 0 || (n > 0 && (!r || !rho || !v || !pr))) { lagb::set_error("sedov_exact_eval: bad argument"); return LAGB_ERR_INVALID; }
      lagb::SedovExact s(dim, gamma, rho0, blast_energy);
      if (alpha_override > 0) { s.set_alpha(alpha_override); }
      s.set_time(t);
      for (int i = 0; i < n; i++) { s.eval(r[i], rho[i], v[i], pr[i]); }
      if (info) { info[0] = s.alpha; info[1] = s.r2; info[2] = s.U; info[3] = s.rho2; info[4] = s.v2; info[5] = s.p2; }
      return LAGB_OK;
   }
   catch (const std::exception &e) { lagb::set_error(e.what()); return LAGB_ERR_INVALID; }
}

int lagb_problem_sedov_density_error(const lagb_problem *p, const double *h_S, const double *h_rho, double t,
                                     double gamma, double rho0, double blast_energy, double out[2])
{
   try
   {
      if (!p || !h_S || !h_rho || !out || !(t > 0)) { lagb::set_error("sedov_density_error: bad argument"); return LAGB_ERR_INVALID; }
      lagb::SedovExact s(p->P.dim, gamma, rho0, blast_energy);
      s.set_time(t);
      out[1] = lagb::sedov_density_error_sum(p->P, h_S, h_rho, s);
      out[0] = std::sqrt(out[1]);
      return LAGB_OK;
   }
   catch (const std::exception &e) { lagb::set_error(e.what()); return LAGB_ERR_INVALID; }
}

// ---- output files (host/mesh_writer.hpp) ----
static int writer_rc(bool ok, const std::string &err) { if (!ok) { lagb::set_error(err); return LAGB_ERR_INVALID; } return LAGB_OK; }

int lagb_problem_write_mesh(const lagb_problem *p, const double *h_x, const char *path, int precision)
{
   if (!p || !path) { lagb::set_error("write_mesh: null argument"); return LAGB_ERR_INVALID; }
   std::string err;
   return writer_rc(lagb::write_text_file(path, err, [&](std::ostream &os)
   { lagb::write_mfem_mesh(os, p->P, h_x, precision > 0 ? precision : 8); }), err);
}

int lagb_problem_write_field(const lagb_problem *p, int kind, int vdim, const double *h_f, const char *path, int precision)
{
   if (!p || !path || !h_f) { lagb::set_error("write_field: null argument"); return LAGB_ERR_INVALID; }
   if (kind != 0 && kind != 1) { lagb::set_error("write_field: kind must be 0 (H1) or 1 (L2)"); return LAGB_ERR_INVALID; }
   if (kind == 0 && vdim < 1) { lagb::set_error("write_field: vdim must be positive"); return LAGB_ERR_INVALID; }
   if (kind == 1 && vdim != 1) { lagb::set_error("write_field: L2 fields are scalar"); return LAGB_ERR_INVALID; }
   const int prec = precision > 0 ? precision : 8;
   std::string err;
   return writer_rc(lagb::write_text_file(path, err, [&](std::ostream &os)
   {
      if (kind == 0) { lagb::write_h1_field(os, p->P, h_f, vdim, prec); } else { lagb::write_l2_field(os, p->P, h_f, prec); }
   }), err);
}

int lagb_problem_write_print(const lagb_problem *p, const char *basename, int ti, const double *h_S,
                             const double *h_rho, int precision)
{
   if (!p || !basename || !h_S || !h_rho) { lagb::set_error("write_print: null argument"); return LAGB_ERR_INVALID; }
   std::string err;
   return writer_rc(lagb::write_print_files(p->P, basename, ti, h_S, h_rho, precision > 0 ? precision : 8, "", err), err);
}

int lagb_problem_write_visit(const lagb_problem *p, const char *collection, int cycle, double time, double time_step,
                             int rank, int nranks, const double *h_S, const double *h_rho, int precision)
{
   if (!p || !collection || !h_S) { lagb::set_error("write_visit: null argument"); return LAGB_ERR_INVALID; }
   if (rank < 0 || nranks < 1 || rank >= nranks || cycle < 0) { lagb::set_error("write_visit: bad rank / cycle"); return LAGB_ERR_INVALID; }
   std::string err;
   return writer_rc(lagb::write_visit_files(p->P, collection, cycle, time, time_step, rank, nranks, h_S, h_rho,
                                            precision > 0 ? precision : 8, err), err);
}

} // extern "C"
