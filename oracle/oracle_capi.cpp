// TEST INFRASTRUCTURE — CPU oracle, not a product path (see oracle/README.md).
// ctypes-facing C API of the oracle: used only by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs.
#include "laghos_oracle.hpp"
#include "../laghos_b200/csrc/host/partition.hpp"

struct OrcHandle
{
   lagb::Problem P;
   oracle::Hydro *H = nullptr;
   ~OrcHandle() { delete H; }
};

extern "C" {

void *orc_create(const char *mesh, int rs, int problem, int ok, int ot, int oq, double blast_scale,
                 int impose_visc, double cfl, double cgt, int cgm, int nthreads)
{
   try
   {
      std::vector<double> coarse[3]; int dim = 0;
      if (!lagb::named_coarse_mesh(mesh, dim, coarse)) { return nullptr; }
      lagb::RectMesh rm; rm.build(dim, coarse, rs);
      lagb::ProblemSpec sp;
      sp.problem = problem; sp.dim = dim; sp.ok = ok; sp.ot = ot; sp.oq = oq;
      sp.blast_scale = blast_scale; sp.impose_visc = impose_visc != 0;
      OrcHandle *h = new OrcHandle();
      h->P.build(sp, rm); oracle::install_own_tables(h->P);
      h->H = new oracle::Hydro(h->P, cfl, cgt, cgm, nthreads);
      return h;
   }
   catch (const std::exception &e) { fprintf(stderr, "orc_create: %s\n", e.what()); return nullptr; }
}
// one rank's element box of a Cartesian partition (tests of the multi-rank host logic)
void *orc_create_part(const char *mesh, int rs, int problem, int ok, int ot, int oq, double blast_scale,
                      int impose_visc, double cfl, double cgt, int cgm, int nthreads, int rank, const int *pgrid)
{
   try
   {
      std::vector<double> coarse[3]; int dim = 0;
      if (!lagb::named_coarse_mesh(mesh, dim, coarse)) { return nullptr; }
      lagb::RectMesh rm; rm.build(dim, coarse, rs);
      lagb::ProblemSpec sp;
      sp.problem = problem; sp.dim = dim; sp.ok = ok; sp.ot = ot; sp.oq = oq;
      sp.blast_scale = blast_scale; sp.impose_visc = impose_visc != 0;
      lagb::Partition part; part.build(dim, rm.n, pgrid, rank, ok);
      OrcHandle *h = new OrcHandle();
      h->P.build(sp, rm, part.lo, part.hi); oracle::install_own_tables(h->P);
      h->H = new oracle::Hydro(h->P, cfl, cgt, cgm, nthreads);
      return h;
   }
   catch (const std::exception &e) { fprintf(stderr, "orc_create_part: %s\n", e.what()); return nullptr; }
}
void orc_destroy(void *h) { delete (OrcHandle*)h; }

// info: dim, NE, D1D, L1D, Q1D, ND, NL, NQ, ndofs_h1, ndofs_l2
void orc_info(void *hh, long long *o)
{
   const lagb::Problem &P = ((OrcHandle*)hh)->P;
   o[0] = P.dim; o[1] = P.NE; o[2] = P.D1D; o[3] = P.L1D; o[4] = P.Q1D; o[5] = P.ND; o[6] = P.NL; o[7] = P.NQ;
   o[8] = P.ndofs_h1; o[9] = P.ndofs_l2;
}
void orc_get_S0(void *hh, double *S) { const auto &v = ((OrcHandle*)hh)->P.S0; memcpy(S, v.data(), sizeof(double)*v.size()); }
double orc_h0(void *hh) { return ((OrcHandle*)hh)->H->qd.h0; }
// which: 0 stressJinvT, 1 rho0DetJ0w, 2 Jac0inv, 3 mass D, 4 mass diagonal
long long orc_qdata(void *hh, int which, double *out)
{
   oracle::Hydro &H = *((OrcHandle*)hh)->H;
   const std::vector<double> *v = nullptr;
   switch (which)
   {
      case 0: v = &H.qd.stressJinvT; break; case 1: v = &H.qd.rho0DetJ0w; break; case 2: v = &H.qd.Jac0inv; break;
      case 3: v = &H.massD; break; case 4: v = &H.diag; break;
   }
   if (!v) { return 0; }
   if (out) { memcpy(out, v->data(), sizeof(double)*v->size()); }
   return (long long)v->size();
}
void orc_set_sjit(void *hh, const double *in)
{
   oracle::Hydro &H = *((OrcHandle*)hh)->H;
   memcpy(H.qd.stressJinvT.data(), in, sizeof(double)*H.qd.stressJinvT.size());
}
void orc_vmass_mult(void *hh, int comp, const double *x, double *y)
{
   OrcHandle *h = (OrcHandle*)hh;
   h->H->VMassMult(comp >= 0 ? &h->P.ess[comp] : nullptr, x, y);
}
void orc_emass_mult(void *hh, const double *x, double *y) { ((OrcHandle*)hh)->H->EMassMult(x, y); }
void orc_force_mult(void *hh, const double *e, double *v) { ((OrcHandle*)hh)->H->ForceMult(e, v); }
void orc_force_mult_t(void *hh, const double *v, double *e) { ((OrcHandle*)hh)->H->ForceMultTranspose(v, e); }
double orc_qupdate(void *hh, const double *S, double dt_in)
{
   oracle::Hydro &H = *((OrcHandle*)hh)->H;
   H.qd.dt_est = dt_in; H.qdata_is_current = false;
   H.UpdateQuadratureData(S);
   H.qdata_is_current = false;
   return H.qd.dt_est;
}
int orc_pcg_vmass(void *hh, int comp, const double *b, double *x)
{
   OrcHandle *h = (OrcHandle*)hh; oracle::Hydro &H = *h->H;
   const std::vector<int> &ess = h->P.ess[comp];
   std::vector<double> B(b, b + h->P.ndofs_h1);
   for (int i : ess) { B[i] = 0.0; }
   return H.CG([&](const double *xx, double *yy) { H.VMassMult(&ess, xx, yy); }, H.dinv.data(), true,
               B.data(), x, h->P.ndofs_h1, H.cg_r.data(), H.cg_d.data(), H.cg_z.data());
}
int orc_cg_emass(void *hh, const double *b, double *x)
{
   OrcHandle *h = (OrcHandle*)hh; oracle::Hydro &H = *h->H;
   return H.CG([&](const double *xx, double *yy) { H.EMassMult(xx, yy); }, nullptr, false,
               b, x, h->P.ndofs_l2, H.l2_r.data(), H.l2_d.data(), H.l2_z.data());
}
void orc_taylor_source(void *hh, const double *x, double *esrc)
{
   OrcHandle *h = (OrcHandle*)hh;
   h->H->K.TaylorSource(h->P, x, esrc);
}
double orc_internal_energy(void *hh, const double *e) { return ((OrcHandle*)hh)->H->InternalEnergy(e); }
double orc_kinetic_energy(void *hh, const double *v) { return ((OrcHandle*)hh)->H->KineticEnergy(v); }
void orc_compute_density(void *hh, const double *x, double *rho) { ((OrcHandle*)hh)->H->ComputeDensity(x, rho); }
// dS_dt = f(S): one call of LagrangianHydroOperator::Mult with fresh quadrature data
void orc_mult(void *hh, const double *S, double *dS_dt)
{
   oracle::Hydro &H = *((OrcHandle*)hh)->H;
   H.ResetTimeStepEstimate(); H.ResetQuadratureData();
   H.Mult(S, dS_dt);
}

// full run; out: [steps, ti_last, t, dt, e_norm, fom0..4, T_cgH1, T_cgL2, T_force, T_qdata, H1iter, L2iter, quad_tstep, stages]
int orc_run(const char *mesh, int rs, int problem, int ok, int ot, int oq, double blast_scale, int impose_visc,
            int ode_solver_type, double t_final, int max_tsteps, double cfl, double cgt, int cgm, int nthreads,
            double *out, double *hist, int hist_cap, double *S_out)
{
   try
   {
      std::vector<double> coarse[3]; int dim = 0;
      if (!lagb::named_coarse_mesh(mesh, dim, coarse)) { return 1; }
      lagb::RectMesh rm; rm.build(dim, coarse, rs);
      lagb::ProblemSpec sp;
      sp.problem = problem; sp.dim = dim; sp.ok = ok; sp.ot = ot; sp.oq = oq;
      sp.blast_scale = blast_scale; sp.impose_visc = impose_visc != 0;
      lagb::Problem P; P.build(sp, rm); oracle::install_own_tables(P);
      oracle::RunOptions o;
      o.ode_solver_type = ode_solver_type; o.t_final = t_final; o.max_tsteps = max_tsteps;
      o.cfl = cfl; o.cg_tol = cgt; o.cg_max_iter = cgm; o.nthreads = nthreads; o.verbose = false;
      std::vector<double> S;
      oracle::RunResult r = oracle::run(P, o, S_out ? &S : nullptr);
      out[0] = r.steps; out[1] = r.ti_last; out[2] = r.t; out[3] = r.dt; out[4] = r.e_norm;
      for (int i = 0; i < 5; i++) { out[5 + i] = r.fom[i]; }
      out[10] = r.timer.sw_cgH1; out[11] = r.timer.sw_cgL2; out[12] = r.timer.sw_force; out[13] = r.timer.sw_qdata;
      out[14] = (double)r.timer.H1iter; out[15] = (double)r.timer.L2iter; out[16] = (double)r.timer.quad_tstep;
      out[17] = r.stages; out[19] = r.energy_init; out[20] = r.energy_final;
      int n = 0;
      for (auto &h : r.e_norm_history) { if (n < hist_cap) { hist[2*n] = h.first; hist[2*n + 1] = h.second; n++; } }
      out[18] = n;
      if (S_out) { memcpy(S_out, S.data(), sizeof(double)*S.size()); }
      return 0;
   }
   catch (const std::exception &e) { fprintf(stderr, "orc_run: %s\n", e.what()); return 2; }
}

} // extern "C"
