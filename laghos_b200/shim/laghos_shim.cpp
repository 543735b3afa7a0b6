// Implementation of the C++ shim (laghos_shim.hpp) and of lagb_laghos_run, the
// restated reference driver loop (laghos.cpp:706-778, 792-839, 928-937).
#include "laghos_shim.hpp"
#include "../csrc/host/problem.hpp"
#include "../csrc/host/partition.hpp"
#include "../csrc/host/mesh_writer.hpp"
#include "../csrc/host/error_norms.hpp"
#include "../csrc/host/checks.hpp"
#include <chrono>
#include <string>

namespace lagb { void set_error(const std::string &msg); }
#include <cmath>
#include <cstring>



namespace laghos {

namespace hydrodynamics {

void MassPAOperator::Mult(const Vector &x, Vector &y) const
{
   if (is_l2) { LAGHOS_CHECK(lagb_emass_mult(ctx, x.Read(), y.Write())); }
   else { LAGHOS_CHECK(lagb_vmass_mult(ctx, ess_comp, x.Read(), y.Write())); }
}
void MassPAOperator::MultFull(const Vector &x, Vector &y) const
{
   if (is_l2) { LAGHOS_CHECK(lagb_emass_mult(ctx, x.Read(), y.Write())); }
   else { LAGHOS_CHECK(lagb_vmass_mult(ctx, -1, x.Read(), y.Write())); }
}

LagrangianHydroOperator::LagrangianHydroOperator(lagb_ctx *ctx_, const lagb_problem_info &info_, double cfl_,
                                                 double cgt, int cgiter, bool batched)
   : TimeDependentOperator(2*info_.dim*info_.ndofs_h1 + info_.ndofs_l2),
     ctx(ctx_), info(info_), dim(info_.dim), source_type(info_.source),
     H1Vsize(info_.dim*info_.ndofs_h1), L2Vsize(info_.ndofs_l2),
     cfl(cfl_), cg_rel_tol(cgt), cg_max_iter(cgiter), batched_pcg(batched),
     qupdate(ctx_, cfl_)
{
   qdata.ctx = ctx; qdata.h0 = lagb_qdata_h0(ctx);
   ForcePA = new ForcePAOperator(ctx);
   VMassPA = new MassPAOperator(ctx, false, info.ndofs_h1);
   EMassPA = new MassPAOperator(ctx, true, info.ndofs_l2);
   // reference laghos_solver.cpp:264-284
   CG_VMass.SetPreconditionerJacobi();
   CG_VMass.SetOperator(*VMassPA);
   CG_VMass.SetRelTol(cg_rel_tol); CG_VMass.SetAbsTol(0.0);
   CG_VMass.SetMaxIter(cg_max_iter); CG_VMass.SetPrintLevel(-1);
   CG_EMass.SetOperator(*EMassPA);
   CG_EMass.iterative_mode = false;
   CG_EMass.SetRelTol(cg_rel_tol); CG_EMass.SetAbsTol(0.0);
   CG_EMass.SetMaxIter(cg_max_iter); CG_EMass.SetPrintLevel(-1);
   one.SetSize(ctx, L2Vsize); one = 1.0;
   rhs.SetSize(ctx, H1Vsize);
   e_rhs.SetSize(ctx, L2Vsize);
   if (source_type == 1) { e_source.SetSize(ctx, L2Vsize); }
   if (source_type == 2)
   {
      // RTCoefficient (laghos_solver.hpp:221-231): accel = (0,-1); B += M accel_c
      // (laghos_solver.cpp:371-380).  M*accel is constant in time (mass is fixed).
      accel_b.SetSize(ctx, H1Vsize); accel_b = 0.0;
      Vector ac(ctx, info.ndofs_h1), b;
      ac = -1.0;
      b.MakeRef(accel_b, info.ndofs_h1, info.ndofs_h1);
      VMassPA->MultFull(ac, b);
   }
}

LagrangianHydroOperator::~LagrangianHydroOperator()
{
   delete EMassPA; delete VMassPA; delete ForcePA;
}

void LagrangianHydroOperator::Mult(const Vector &S, Vector &dS_dt) const
{
   UpdateMesh(S);
   Vector v, dx;
   v.MakeRef(S, H1Vsize, H1Vsize);
   dx.MakeRef(dS_dt, 0, H1Vsize);
   // reference order is dx = v; SolveVelocity; SolveEnergy (laghos_solver.cpp:316-325).  dx and dv are
   // disjoint slices of dS_dt, so dx = v may follow SolveVelocity: with current quadrature data
   // SolveVelocity does not read S, which lets a background H2D copy of S overlap Force + PCG.
   SolveVelocity(S, dS_dt);
   WaitState();
   dx = v;
   SolveEnergy(S, v, dS_dt);
   qdata_is_current = false;
}

void LagrangianHydroOperator::UpdateQuadratureData(const Vector &S) const
{
   if (qdata_is_current) { return; }
   qdata_is_current = true;
   WaitState();
   qupdate.UpdateQuadratureData(S, qdata);
}

void LagrangianHydroOperator::SolveVelocity(const Vector &S, Vector &dS_dt) const
{
   UpdateQuadratureData(S);
   Vector dv;
   dv.MakeRef(dS_dt, H1Vsize, H1Vsize);
   if (!batched_pcg) { dv = 0.0; }      // the batched solve starts from zero itself (lagb_pcg_vmass_all_x0)
   ForcePA->Mult(one, rhs);
   rhs.Neg();
   if (source_type == 2) { rhs.Add(1.0, accel_b); }
   const int64_t size = info.ndofs_h1;
   if (batched_pcg)
   {
      // all components in one batched device PCG: the quadrature data is streamed
      // once per iteration for the `dim` solves of laghos_solver.cpp:363-398
      int it[3];
      // zero initial guess (dS_dt = 0 in the reference): x = 0, r = b, no operator application before the loop
      LAGHOS_CHECK(lagb_pcg_vmass_all_x0(ctx, rhs.Read(), dv.Write(), cg_rel_tol, cg_max_iter, it));
   }
   else
   {
      for (int c = 0; c < dim; c++)
      {
         Vector dvc, B;
         dvc.MakeRef(dS_dt, H1Vsize + c*size, size);
         B.MakeRef(rhs, c*size, size);
         VMassPA->SetEssentialTrueDofs(c);
         VMassPA->EliminateRHS(B);
         CG_VMass.Mult(B, dvc);
      }
   }
}

void LagrangianHydroOperator::SolveEnergy(const Vector &S, const Vector &v, Vector &dS_dt) const
{
   UpdateQuadratureData(S);
   Vector de;
   de.MakeRef(dS_dt, 2*H1Vsize, L2Vsize);
   de = 0.0;
   if (source_type == 1) { LAGHOS_CHECK(lagb_taylor_source(ctx, S.Read(), e_source.Write())); }
   ForcePA->MultTranspose(v, e_rhs);
   if (source_type == 1) { e_rhs.Add(1.0, e_source); }
   CG_EMass.Mult(e_rhs, de);
}

double LagrangianHydroOperator::GetTimeStepEstimate(const Vector &S) const
{
   UpdateMesh(S);
   UpdateQuadratureData(S);
   double dt;
   LAGHOS_CHECK(lagb_dt_est_read(ctx, &dt));
   qdata.dt_est = dt;
   return dt;
}

void LagrangianHydroOperator::ResetTimeStepEstimate() const
{
   qdata.dt_est = std::numeric_limits<double>::infinity();
   LAGHOS_CHECK(lagb_dt_est_set(ctx, qdata.dt_est));
}

void LagrangianHydroOperator::GetFOM(long long steps, double fom[5], lagb_timing &tm) const
{
   LAGHOS_CHECK(lagb_timing_get(ctx, &tm));
   // global sizes and max-over-ranks times are reduced by the caller for nranks > 1
   const double H1GTVSize = (double)H1Vsize, L2GTVSize = (double)L2Vsize;
   const long long H1iter = tm.H1iter/dim;
   const double T0 = tm.t_cgH1, T2 = tm.t_force, T3 = tm.t_qdata, T4 = T0 + T2 + T3;
   fom[1] = 1e-6*H1GTVSize*H1iter/T0;
   fom[2] = 1e-6*steps*(H1GTVSize + L2GTVSize)/T2;
   fom[3] = 1e-6*tm.quad_tstep*info.NQ/T3;
   fom[0] = (fom[1]*T0 + fom[2]*T2 + fom[3]*T3)/T4;
   fom[4] = T4;
}

void LagrangianHydroOperator::PrintTimingData(bool IamRoot, long long steps, bool) const
{
   double fom[5]; lagb_timing tm;
   GetFOM(steps, fom, tm);
   if (!IamRoot) { return; }
   printf("\nCG (H1) total time: %g\nCG (H1) rate (megadofs x cg_iterations / second): %g\n", tm.t_cgH1, fom[1]);
   printf("\nCG (L2) total time: %g\nCG (L2) rate (megadofs x cg_iterations / second): %g\n", tm.t_cgL2,
          1e-6*(double)L2Vsize*tm.L2iter/tm.t_cgL2);
   printf("\nForces total time: %g\nForces rate (megadofs x timesteps / second): %g\n", tm.t_force, fom[2]);
   printf("\nUpdateQuadData total time: %g\nUpdateQuadData rate (megaquads x timesteps / second): %g\n", tm.t_qdata, fom[3]);
   printf("\nMajor kernels total time (seconds): %g\nMajor kernels total rate (megadofs x time steps / second): %g\n", fom[4], fom[0]);
}

} // namespace hydrodynamics

void CGSolver::SetOperator(const Operator &op)
{
   oper = dynamic_cast<const hydrodynamics::MassPAOperator*>(&op);
   if (!oper) { LAGHOS_ABORT("CGSolver: the device PCG expects a MassPAOperator"); }
   height = width = op.Height();
}

void CGSolver::Mult(const Vector &b, Vector &x) const
{
   int it = 0;
   if (oper->IsL2())
   {
      if (iterative_mode) { LAGHOS_ABORT("CG_EMass runs with iterative_mode = false (laghos_solver.cpp:279)"); }
      LAGHOS_CHECK(lagb_cg_emass(oper->Ctx(), b.Read(), x.Write(), rel_tol, max_iter, &it));
   }
   else
   {
      if (!have_prec || !iterative_mode) { LAGHOS_ABORT("CG_VMass runs Jacobi-preconditioned with iterative_mode = true"); }
      LAGHOS_CHECK(lagb_pcg_vmass(oper->Ctx(), oper->EssComp(), b.Read(), x.Write(), rel_tol, max_iter, &it));
   }
   final_iter = it;
}

// ---- ODE solvers: MFEM ForwardEuler/RK2/RK3SSP/RK4 restated ----
void ForwardEulerSolver::Init(TimeDependentOperator &f_) { ODESolver::Init(f_); }
void ForwardEulerSolver::Step(Vector &x, double &t, double &dt)
{
   if (dxdt.Size() != x.Size()) { dxdt.SetSize(x.Ctx(), x.Size()); }
   f->SetTime(t); f->Mult(x, dxdt); x.Add(dt, dxdt); t += dt;
}
void RK2Solver::Init(TimeDependentOperator &f_) { ODESolver::Init(f_); }
void RK2Solver::Step(Vector &x, double &t, double &dt)
{
   if (dxdt.Size() != x.Size()) { dxdt.SetSize(x.Ctx(), x.Size()); x1.SetSize(x.Ctx(), x.Size()); }
   const double b = 0.5/a;
   f->SetTime(t); f->Mult(x, dxdt);
   add(x, (1. - b)*dt, dxdt, x1);
   x.Add(a*dt, dxdt);
   f->SetTime(t + a*dt); f->Mult(x, dxdt);
   add(x1, b*dt, dxdt, x);
   t += dt;
}
void RK3SSPSolver::Init(TimeDependentOperator &f_) { ODESolver::Init(f_); }
void RK3SSPSolver::Step(Vector &x, double &t, double &dt)
{
   if (y.Size() != x.Size()) { y.SetSize(x.Ctx(), x.Size()); k.SetSize(x.Ctx(), x.Size()); }
   lagb_ctx *c = x.Ctx();
   f->SetTime(t); f->Mult(x, k);
   add(x, dt, k, y);
   f->SetTime(t + dt); f->Mult(y, k);
   y.Add(dt, k);
   LAGHOS_CHECK(lagb_vec_axpby(c, y.Write(), 3./4, x.Read(), 1./4, y.Read(), x.Size()));
   f->SetTime(t + dt/2); f->Mult(y, k);
   y.Add(dt, k);
   LAGHOS_CHECK(lagb_vec_axpby(c, x.Write(), 1./3, x.Read(), 2./3, y.Read(), x.Size()));
   t += dt;
}
void RK4Solver::Init(TimeDependentOperator &f_) { ODESolver::Init(f_); }
void RK4Solver::Step(Vector &x, double &t, double &dt)
{
   if (y.Size() != x.Size()) { y.SetSize(x.Ctx(), x.Size()); k.SetSize(x.Ctx(), x.Size()); z.SetSize(x.Ctx(), x.Size()); }
   f->SetTime(t); f->Mult(x, k);            // k1
   add(x, dt/2, k, y);
   add(x, dt/6, k, z);
   f->SetTime(t + dt/2); f->Mult(y, k);     // k2
   add(x, dt/2, k, y);
   z.Add(dt/3, k);
   f->Mult(y, k);                           // k3
   add(x, dt, k, y);
   z.Add(dt/3, k);
   f->SetTime(t + dt); f->Mult(y, k);       // k4
   add(z, dt/6, k, x);
   t += dt;
}
void RK6Solver::Init(TimeDependentOperator &f_) { ODESolver::Init(f_); }
void RK6Solver::Step(Vector &x, double &t, double &dt)
{
   static const double a[28] = {.6e-1,
   .1923996296296296296296296296296296296296e-1, .7669337037037037037037037037037037037037e-1,
   .35975e-1, 0., .107925,
   1.318683415233148260919747276431735612861, 0., -5.042058063628562225427761634715637693344, 4.220674648395413964508014358283902080483,
   -41.87259166432751461803757780644346812905, 0., 159.4325621631374917700365669070346830453, -122.1192135650100309202516203389242140663, 5.531743066200053768252631238332999150076,
   -54.43015693531650433250642051294142461271, 0., 207.0672513650184644273657173866509835987, -158.6108137845899991828742424365058599469, 6.991816585950242321992597280791793907096, -.1859723106220323397765171799549294623692e-1,
   -54.66374178728197680241215648050386959351, 0., 207.9528062553893734515824816699834244238, -159.2889574744995071508959805871426654216, 7.018743740796944434698170760964252490817, -.1833878590504572306472782005141738268361e-1, -.5119484997882099077875432497245168395840e-3};
   static const double b[8] = {.3438957868357036009278820124728322386520e-1, 0., 0., .2582624555633503404659558098586120858767, .4209371189673537150642551514069801967032,
   4.405396469669310170148836816197095664891, -176.4831190242986576151740942499002125029, 172.3641334014150730294022582711902413315};
   static const double c[7] = {.6e-1, .9593333333333333333333333333333333333333e-1, .1439, .4973, .9725, .9995, 1.};
   if (y.Size() != x.Size()) { y.SetSize(x.Ctx(), x.Size()); for (auto &ki : k) { ki.SetSize(x.Ctx(), x.Size()); } }
   f->SetTime(t); f->Mult(x, k[0]);
   for (int l = 0, i = 1; i < 8; i++)
   {
      add(x, a[l++]*dt, k[0], y);
      for (int j = 1; j < i; j++) { y.Add(a[l++]*dt, k[j]); }
      f->SetTime(t + c[i - 1]*dt); f->Mult(y, k[i]);
   }
   for (int i = 0; i < 8; i++) { x.Add(b[i]*dt, k[i]); }
   t += dt;
}
void HydroODESolver::Init(TimeDependentOperator &f_)
{
   ODESolver::Init(f_);
   hydro_oper = dynamic_cast<hydrodynamics::LagrangianHydroOperator*>(f);
   if (!hydro_oper) { LAGHOS_ABORT("HydroSolvers expect LagrangianHydroOperator."); }
}
void RK2AvgSolver::Init(TimeDependentOperator &f_) { HydroODESolver::Init(f_); }
void RK2AvgSolver::Step(Vector &S, double &t, double &dt)
{
   const int64_t NV = hydro_oper->GetH1VSize();
   if (S0.Size() != S.Size())
   {
      S0.SetSize(S.Ctx(), S.Size()); dS_dt.SetSize(S.Ctx(), S.Size()); V.SetSize(S.Ctx(), NV);
      dS_dt = 0.0;
   }
   hydro_oper->WaitState();   // S may still be arriving from the host (bench e2e)
   S0 = S;
   Vector v0, dx_dt, dv_dt;
   v0.MakeRef(S0, NV, NV); dx_dt.MakeRef(dS_dt, 0, NV); dv_dt.MakeRef(dS_dt, NV, NV);
   hydro_oper->UpdateMesh(S);
   hydro_oper->SolveVelocity(S, dS_dt);
   add(v0, 0.5*dt, dv_dt, V);
   hydro_oper->SolveEnergy(S, V, dS_dt);
   dx_dt = V;
   add(S0, 0.5*dt, dS_dt, S);
   hydro_oper->ResetQuadratureData();
   hydro_oper->UpdateMesh(S);
   hydro_oper->SolveVelocity(S, dS_dt);
   add(v0, 0.5*dt, dv_dt, V);
   hydro_oper->SolveEnergy(S, V, dS_dt);
   dx_dt = V;
   add(S0, dt, dS_dt, S);
   hydro_oper->ResetQuadratureData();
   t += dt;
}

} // namespace laghos

using namespace laghos;

extern "C" void lagb_run_options_default(lagb_run_options *o)
{
   memset(o, 0, sizeof(*o));
   o->mesh = "cube01_hex"; o->rs = 2; o->problem = 1; o->ok = 2; o->ot = 1; o->oq = -1;
   o->blast_scale = 0.125; o->ode_solver_type = 4; o->t_final = 0.6; o->max_tsteps = -1;
   o->cfl = 0.5; o->cg_tol = 1e-8; o->cg_max_iter = 300; o->batched_pcg = 1; o->vis_steps = 5;
   o->nranks = 1; o->pgrid[0] = o->pgrid[1] = o->pgrid[2] = 1;
   o->basename = "results/Laghos";   // laghos.cpp:170
}

extern "C" int lagb_laghos_run(const lagb_run_options *opt, lagb_run_result *res, double *hist, int hist_cap, double *S_out)
{
   memset(res, 0, sizeof(*res));
   std::vector<double> coarse[3]; int dim = 0;
   if (!opt->mesh || !lagb::named_coarse_mesh(opt->mesh, dim, coarse)) { fprintf(stderr, "unknown mesh\n"); return LAGB_ERR_INVALID; }
   lagb::RectMesh gm; gm.build(dim, coarse, opt->rs);
   lagb::ProblemSpec sp;
   sp.problem = opt->problem; sp.dim = dim; sp.ok = opt->ok; sp.ot = opt->ot; sp.oq = opt->oq;
   sp.blast_scale = opt->blast_scale; sp.impose_visc = opt->impose_visc != 0;
   lagb::Problem P;
   lagb::Partition part;
   const int nranks = std::max(1, opt->nranks);
   try
   {
      if (nranks > 1)
      {
         part.build(dim, gm.n, opt->pgrid, opt->rank, opt->ok);
         if (part.nranks != nranks) { fprintf(stderr, "pgrid does not match nranks\n"); return LAGB_ERR_INVALID; }
         P.build(sp, gm, part.lo, part.hi);
      }
      else { P.build(sp, gm); }
   }
   catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return LAGB_ERR_INVALID; }

   lagb_ctx_desc d; memset(&d, 0, sizeof d);
   d.dim = P.dim; d.NE = P.NE; d.D1D = P.D1D; d.L1D = P.L1D; d.Q1D = P.Q1D; d.ndofs_h1 = P.ndofs_h1;
   d.h_h1_map = P.h1_map.data();
   for (int c = 0; c < P.dim; c++) { d.h_ess[c] = P.ess[c].data(); d.ness[c] = (int)P.ess[c].size(); }
   d.h_B = P.tab.B.data(); d.h_G = P.tab.G.data(); d.h_BL = P.tab.BL.data();
   d.h_qweights = P.qweights.data(); d.h_gamma = P.gamma.data();
   d.use_visc = P.use_visc; d.use_vort = P.use_vort; d.device = opt->device; d.kernel_variant = opt->kernel_variant;
   for (int k = 0; k < 3; k++) { d.elem_grid[k] = P.nloc[k]; }   // Cartesian block, lexicographic element numbering
   lagb_ctx *ctx = nullptr;
   LAGHOS_CHECK(lagb_ctx_create(&ctx, &d, nullptr));
   if (nranks > 1)
   {
      std::vector<int32_t> nr, ph, ns; std::vector<const int32_t*> lists;
      for (auto &nb : part.nbrs) { nr.push_back(nb.rank); ph.push_back(nb.phase); ns.push_back((int)nb.dofs.size()); lists.push_back(nb.dofs.data()); }
      LAGHOS_CHECK(lagb_ctx_comm_init(ctx, opt->nccl_id, opt->rank, nranks, (int)nr.size(), nr.data(), ph.data(), ns.data(),
                                      lists.data(), part.owner.data()));
   }
   lagb_problem_info info; memset(&info, 0, sizeof info);
   info.dim = P.dim; info.NE = P.NE; info.D1D = P.D1D; info.L1D = P.L1D; info.Q1D = P.Q1D;
   info.ND = P.ND; info.NL = P.NL; info.NQ = P.NQ; info.ndofs_h1 = P.ndofs_h1; info.ndofs_l2 = P.ndofs_l2;
   info.use_visc = P.use_visc; info.use_vort = P.use_vort; info.source = P.source;

   int rc = 0;
   {
      const int64_t N = P.s_size(), NV = P.h1_vsize();
      Vector S(ctx, N), S_old(ctx, N), rho0(ctx, P.ndofs_l2), rho0q(ctx, (int64_t)P.NE*P.NQ);
      double *S_pin = nullptr;
      if (opt->e2e_host_state)
      {
         LAGHOS_CHECK(lagb_host_alloc_pinned(&S_pin, N));
         memcpy(S_pin, P.S0.data(), sizeof(double)*N);
      }
      S.HostWrite(P.S0.data()); rho0.HostWrite(P.rho0_gf.data()); rho0q.HostWrite(P.rho0_q.data());
      double h0;
      LAGHOS_CHECK(lagb_setup_qdata0(ctx, S.Read(), rho0.Read(), rho0q.Read(), gm.NE(), &h0));

      hydrodynamics::LagrangianHydroOperator hydro(ctx, info, opt->cfl, opt->cg_tol, opt->cg_max_iter, opt->batched_pcg != 0);
      ODESolver *ode_solver = nullptr;
      int stages = 1;
      switch (opt->ode_solver_type)   // reference laghos.cpp:519-534
      {
         case 1: ode_solver = new ForwardEulerSolver; break;
         case 2: ode_solver = new RK2Solver(0.5); stages = 2; break;
         case 3: ode_solver = new RK3SSPSolver; stages = 3; break;
         case 4: ode_solver = new RK4Solver; stages = 4; break;
         case 6: ode_solver = new RK6Solver; stages = 8; break;
         case 7: ode_solver = new RK2AvgSolver; stages = 2; break;
         default: fprintf(stderr, "Unknown ODE solver type: %d\n", opt->ode_solver_type); lagb_ctx_destroy(ctx); return 3;
      }
      ode_solver->Init(hydro);
      Vector v_gf0, e_gf0; v_gf0.MakeRef(S, NV, NV); e_gf0.MakeRef(S, 2*NV, P.ndofs_l2);
      res->energy_init = hydro.InternalEnergy(e_gf0) + hydro.KineticEnergy(v_gf0);   // laghos.cpp:664-665
      hydro.ResetTimeStepEstimate();
      double t = 0.0, dt = hydro.GetTimeStepEstimate(S), t_old;
      bool last_step = false;
      int steps = 0, timed_from_step = 0;
      bool sw_started = false;
      Vector e_gf; e_gf.MakeRef(S, 2*NV, P.ndofs_l2);
      auto e_norm = [&]()
      {
         double lnorm = e_gf*e_gf;
         if (nranks > 1) { LAGHOS_CHECK(lagb_allreduce_host(ctx, &lnorm, 1, 0)); }
         return std::sqrt(lnorm);
      };
      auto t_wall0 = std::chrono::steady_clock::now();
      int64_t launches0 = lagb_kernel_launch_count();
      LAGHOS_CHECK(lagb_profile_mass(ctx, opt->profile_mass));
      if (opt->warmup_steps <= 0)
      {
         LAGHOS_CHECK(lagb_timing_reset(ctx)); LAGHOS_CHECK(lagb_ctx_sync(ctx));
         t_wall0 = std::chrono::steady_clock::now();
         LAGHOS_CHECK(lagb_stopwatch_start(ctx)); sw_started = true;
      }
      int n_hist = 0, ti = 1, checks = 0;
      if (opt->check)
      {
         const std::string bad = lagb::checks_preconditions(opt->mesh, P.dim, opt->rs, opt->ok, opt->ot, opt->ode_solver_type,
                                                            opt->t_final, opt->cfl);
         if (!bad.empty()) { lagb::set_error(bad); rc = LAGB_ERR_INVALID; last_step = true; }
      }
      for (; !last_step; ti++)
      {
         if (t + dt >= opt->t_final) { dt = opt->t_final - t; last_step = true; }
         if (steps == opt->max_tsteps) { last_step = true; }
         if (opt->e2e_host_state)
         {
            // S comes from pinned host memory every step; the copy runs on the copy stream behind the
            // previous step's D2H and overlaps the first Force + PCG of the step (which do not read S).
            // The host copy doubles as the rollback state (it is only overwritten by accepted steps).
            LAGHOS_CHECK(lagb_memcpy_h2d_bg(ctx, S.Write(), S_pin, N));
            hydro.StatePending();
         }
         else { S_old = S; }
         t_old = t;
         hydro.ResetTimeStepEstimate();
         ode_solver->Step(S, t, dt);
         steps++;
         const double dt_est = hydro.GetTimeStepEstimate(S);
         if (dt_est < dt)
         {
            dt *= 0.85;
            if (dt < std::numeric_limits<double>::epsilon()) { LAGHOS_ABORT("The time step crashed!"); }
            t = t_old;
            if (!opt->e2e_host_state) { S = S_old; }   // e2e: the next iteration reloads S from the host copy
            hydro.ResetQuadratureData();
            if (opt->verbose && opt->rank == 0) { printf("Repeating step %d\n", ti); }
            if (steps < opt->max_tsteps) { last_step = false; }
            ti--;
         }
         else
         {
            if (dt_est > 1.25*dt) { dt *= 1.02; }
            if (opt->e2e_host_state) { LAGHOS_CHECK(lagb_memcpy_d2h_bg(ctx, S_pin, S.Read(), N)); }   // overlaps the next step's start
            const bool print = opt->verbose && (last_step || (ti % std::max(1, opt->vis_steps)) == 0);
            if (hist_cap > 0 || print || last_step || opt->check)
            {
               const double nrm = e_norm();
               res->e_norm = nrm; res->ti_last = ti;
               if (hist && n_hist < hist_cap) { hist[2*n_hist] = ti; hist[2*n_hist + 1] = nrm; n_hist++; }
               if (print && opt->rank == 0) { printf("step %5d,\tt = %5.4f,\tdt = %5.6f,\t|e| = %.10e\n", ti, t, dt, nrm); }
               if (opt->check)
               {
                  std::string msg;
                  if (lagb::checks_step(P.dim, opt->problem, ti, nrm, opt->check_eps > 0 ? opt->check_eps : 1e-13, checks, msg) < 0)
                  {
                     // the reference aborts here (MFEM_VERIFY); a library call reports the failure instead: the loop
                     // stops, the context is released below and the status + message go back to the caller
                     if (opt->rank == 0) { printf("%.15e\n", nrm); }
                     lagb::set_error("check failed: " + msg);
                     rc = LAGB_ERR_INVALID; last_step = true;
                  }
               }
            }
            if ((opt->gfprint || opt->visit) && (last_step || (ti % std::max(1, opt->vis_steps)) == 0))
            {
               // laghos.cpp:845-900: density projected on the current mesh, then mesh + rho + v + e to files
               Vector x_now, rho_gf(ctx, P.ndofs_l2);
               x_now.MakeRef(S, 0, NV);
               hydro.ComputeDensity(x_now, rho_gf);
               std::vector<double> hS, hrho;
               S.HostRead(hS); rho_gf.HostRead(hrho);
               const std::string base = opt->basename ? opt->basename : "results/Laghos";
               std::string err; char sfx[16] = "";
               if (nranks > 1) { snprintf(sfx, sizeof sfx, ".%06d", opt->rank); }
               bool ok = true;
               if (opt->gfprint) { ok = lagb::write_print_files(P, base, ti, hS.data(), hrho.data(), 8, sfx, err); }
               if (ok && opt->visit) { ok = lagb::write_visit_files(P, base, ti, t, dt, opt->rank, nranks, hS.data(), hrho.data(), 8, err); }
               if (!ok) { LAGHOS_ABORT(err.c_str()); }
            }
         }
         if (opt->warmup_steps > 0 && steps == opt->warmup_steps)
         {
            LAGHOS_CHECK(lagb_timing_reset(ctx)); LAGHOS_CHECK(lagb_ctx_sync(ctx));
            t_wall0 = std::chrono::steady_clock::now(); launches0 = lagb_kernel_launch_count();
            LAGHOS_CHECK(lagb_stopwatch_start(ctx)); sw_started = true;
            timed_from_step = steps;
         }
      }
      if (!sw_started) { LAGHOS_CHECK(lagb_stopwatch_start(ctx)); }
      if (opt->e2e_host_state) { LAGHOS_CHECK(lagb_wait_copies(ctx)); }   // the timed region ends after the last D2H
      LAGHOS_CHECK(lagb_stopwatch_stop(ctx, &res->device_seconds));
      const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_wall0).count();
      LAGHOS_CHECK(lagb_profile_mass_get(ctx, &res->mass_kernel_seconds, &res->mass_kernel_launches));
      res->mass_kernel_ncomp = opt->batched_pcg ? P.dim : 1;
      res->steps = steps; res->t = t; res->dt = dt; res->stages = stages; res->n_hist = n_hist;
      res->checks = checks;
      if (opt->check && rc == 0 && checks != 2) { lagb::set_error("Check error!"); rc = LAGB_ERR_INVALID; }   // laghos.cpp:926
      res->wall_seconds = wall;
      res->kernel_launches = lagb_kernel_launch_count() - launches0;
      res->h2d_bytes_per_step = opt->e2e_host_state ? (int64_t)N*8 : 0;
      res->d2h_bytes_per_step = opt->e2e_host_state ? (int64_t)N*8 + 8 : 8;
      const long long timed_steps = (long long)(steps - timed_from_step)*stages;
      // global sizes (duplicates on shared faces are counted once via the owner mask),
      // max-over-ranks timers and summed counters: reference laghos_solver.cpp:703-727
      double sizes[4] = {(double)P.NE, (double)P.ndofs_l2, 0.0, 0.0};
      if (nranks > 1) { for (unsigned char o : part.owner) { sizes[2] += o; } }
      else { sizes[2] = (double)P.ndofs_h1; }
      LAGHOS_CHECK(lagb_timing_get(ctx, &res->timing));
      sizes[3] = (double)res->timing.quad_tstep;
      double times[4] = {res->timing.t_cgH1, res->timing.t_cgL2, res->timing.t_force, res->timing.t_qdata};
      if (nranks > 1)
      {
         LAGHOS_CHECK(lagb_allreduce_host(ctx, sizes, 4, 0));
         LAGHOS_CHECK(lagb_allreduce_host(ctx, times, 4, 2));
      }
      res->ne_global = (int64_t)sizes[0]; res->ndofs_l2_global = (int64_t)sizes[1]; res->ndofs_h1_global = (int64_t)sizes[2];
      res->timing.t_cgH1 = times[0]; res->timing.t_cgL2 = times[1]; res->timing.t_force = times[2]; res->timing.t_qdata = times[3];
      res->timing.quad_tstep = (int64_t)sizes[3];
      {
         const double H1GTVSize = (double)P.dim*sizes[2], L2GTVSize = sizes[1];
         const long long H1iter = res->timing.H1iter/P.dim;
         const double T0 = times[0], T2 = times[2], T3 = times[3], T4 = T0 + T2 + T3;
         res->fom[1] = 1e-6*H1GTVSize*H1iter/T0;
         res->fom[2] = 1e-6*timed_steps*(H1GTVSize + L2GTVSize)/T2;
         res->fom[3] = 1e-6*sizes[3]*P.NQ/T3;
         res->fom[0] = (res->fom[1]*T0 + res->fom[2]*T2 + res->fom[3]*T3)/T4;
         res->fom[4] = T4;
         res->work_mdof = 1e-6*(H1GTVSize*H1iter + timed_steps*(H1GTVSize + L2GTVSize) + sizes[3]*P.NQ);
      }
      res->energy_final = hydro.InternalEnergy(e_gf0) + hydro.KineticEnergy(v_gf0);      // laghos.cpp:956-962
      if (opt->verbose && opt->rank == 0) { printf("\nEnergy  diff: %.2e\n", std::fabs(res->energy_init - res->energy_final)); }
      if (S_out) { std::vector<double> tmp; S.HostRead(tmp); memcpy(S_out, tmp.data(), sizeof(double)*N); }
      if (opt->check_exact_sedov)
      {
         // -err (laghos.cpp:1009-1085): density projected on the final mesh against the exact Sedov solution at time t
         // (gamma = 1.4, rho0 = 1, omega = 0, blast at the origin)
         if (opt->problem != 1) { LAGHOS_ABORT("Can only compare problem 1 (Sedov) against the exact solution"); }
         lagb::SedovExact asol(P.dim, 1.4, 1.0, opt->blast_scale*(1 << P.dim));
         asol.set_time(t);
         double min_r = 1e300;
         for (int dd = 0; dd < P.dim; dd++) { min_r = std::min(min_r, gm.brk[dd].back()); }
         if (!(asol.r2 <= min_r)) { LAGHOS_ABORT("Solution reflections off boundaries detected, cannot compare against exact solution."); }
         Vector x_now, rho_gf(ctx, P.ndofs_l2);
         x_now.MakeRef(S, 0, NV);
         hydro.ComputeDensity(x_now, rho_gf);
         std::vector<double> hS, hrho;
         S.HostRead(hS); rho_gf.HostRead(hrho);
         double sum = lagb::sedov_density_error_sum(P, hS.data(), hrho.data(), asol);
         if (nranks > 1) { LAGHOS_CHECK(lagb_allreduce_host(ctx, &sum, 1, 0)); }
         res->density_l2_err = std::sqrt(sum);
         if (opt->verbose && opt->rank == 0) { printf("Density L2 error: %g\n", res->density_l2_err); }
      }
      if (opt->v_error && (opt->problem == 0 || opt->problem == 4))
      {
         // laghos.cpp:970-982: for problems 0 and 4 the exact velocity is constant in time
         std::vector<double> hS; S.HostRead(hS);
         double s[3];
         lagb::velocity_error_sums(P, hS.data(), s);
         if (nranks > 1) { LAGHOS_CHECK(lagb_allreduce_host(ctx, &s[0], 1, 2)); LAGHOS_CHECK(lagb_allreduce_host(ctx, &s[1], 2, 0)); }
         res->v_err[0] = s[0]; res->v_err[1] = s[1]; res->v_err[2] = std::sqrt(s[2]);
         if (opt->verbose && opt->rank == 0)
         { printf("L_inf  error: %g\nL_1    error: %g\nL_2    error: %g\n", res->v_err[0], res->v_err[1], res->v_err[2]); }
      }
      if (opt->verbose && opt->rank == 0)
      {
         printf("\nCG (H1) total time: %g\nCG (H1) rate (megadofs x cg_iterations / second): %g\n", times[0], res->fom[1]);
         printf("CG (L2) total time: %g\n", times[1]);
         printf("Forces total time: %g\nForces rate (megadofs x timesteps / second): %g\n", times[2], res->fom[2]);
         printf("UpdateQuadData total time: %g\nUpdateQuadData rate (megaquads x timesteps / second): %g\n", times[3], res->fom[3]);
         printf("Major kernels total time (seconds): %g\nMajor kernels total rate (megadofs x time steps / second): %g\n", res->fom[4], res->fom[0]);
      }
      delete ode_solver;
      if (S_pin) { lagb_host_free_pinned(S_pin); }
   }
   lagb_ctx_destroy(ctx);
   return rc;
}
