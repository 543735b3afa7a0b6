// C++ host shim over the C ABI (include/laghos_b200.h).
//
// Keeps the reference's class and method names for the hot path so that the
// reference's time loop (laghos.cpp:706-778) instantiates these operators
// unchanged in structure:
//   Vector, Operator, Solver, CGSolver, TimeDependentOperator, ODESolver,
//   RK4Solver (MFEM, not in tree), RK2AvgSolver (laghos_solver.hpp:246-255),
//   hydrodynamics::{QuadratureData, TimingData, MassPAOperator, ForcePAOperator,
//   QUpdate, LagrangianHydroOperator} (laghos_assembly.hpp:31-131,
//   laghos_solver.hpp:39-205).
// Differences that are deliberate (DESIGN.md "boundary"):
//   * Vector owns DEVICE memory only (MFEM's Memory<> validity flags are replaced by
//     explicit residency); host access goes through HostRead()/HostWrite() copies.
//   * Errors: a non-zero C-ABI status aborts with the library's message, like
//     MFEM_ABORT / MFEM_VERIFY (no exceptions, no error returns).
//   * CGSolver::Mult forwards to the fused device-resident PCG
//     (lagb_pcg_vmass / lagb_cg_emass) when its operator is a MassPAOperator.
#pragma once
#include "../../include/laghos_b200.h"
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>

namespace laghos {

#define LAGHOS_ABORT(msg) do { fprintf(stderr, "\nLAGHOS abort: %s\n ... in %s:%d\n", msg, __FILE__, __LINE__); abort(); } while (0)
#define LAGHOS_CHECK(call) do { if ((call) != 0) { LAGHOS_ABORT(lagb_last_error()); } } while (0)

class Vector
{
   lagb_ctx *ctx = nullptr;
   double *d = nullptr;
   int64_t n = 0;
   bool own = false;
public:
   Vector() { }
   Vector(lagb_ctx *c, int64_t size) { SetSize(c, size); }
   Vector(const Vector &) = delete;
   Vector &operator=(const Vector &v) { LAGHOS_CHECK(lagb_vec_copy(ctx, d, v.d, n)); return *this; }
   ~Vector() { Destroy(); }
   void Destroy() { if (own && d) { lagb_dev_free(ctx, d); } d = nullptr; n = 0; own = false; }
   void SetSize(lagb_ctx *c, int64_t size) { Destroy(); ctx = c; n = size; own = true; LAGHOS_CHECK(lagb_dev_malloc(c, &d, size)); }
   // alias of a sub-range of base (MFEM Vector::MakeRef / GridFunction::MakeRef)
   void MakeRef(const Vector &base, int64_t offset, int64_t size) { Destroy(); ctx = base.ctx; d = base.d + offset; n = size; own = false; }
   int64_t Size() const { return n; }
   lagb_ctx *Ctx() const { return ctx; }
   const double *Read() const { return d; }    // device pointer, as mfem::Vector::Read() on a device build
   double *Write() { return d; }
   double *ReadWrite() { return d; }
   Vector &operator=(double a) { LAGHOS_CHECK(lagb_vec_fill(ctx, d, a, n)); return *this; }
   void Neg() { LAGHOS_CHECK(lagb_vec_axpby(ctx, d, -1.0, d, 0.0, nullptr, n)); }
   Vector &Add(double a, const Vector &x) { LAGHOS_CHECK(lagb_vec_axpby(ctx, d, 1.0, d, a, x.d, n)); return *this; } // *this += a x
   double operator*(const Vector &y) const { double r; LAGHOS_CHECK(lagb_vec_dot(ctx, d, y.d, n, &r)); return r; }
   void HostRead(std::vector<double> &h) const { h.resize(n); LAGHOS_CHECK(lagb_memcpy_d2h(ctx, h.data(), d, n)); }
   void HostWrite(const double *h) { LAGHOS_CHECK(lagb_memcpy_h2d(ctx, d, h, n)); }
};
// z = x + a*y  (mfem::add(x, a, y, z))
inline void add(const Vector &x, double a, const Vector &y, Vector &z)
{ LAGHOS_CHECK(lagb_vec_axpby(z.Ctx(), z.Write(), 1.0, x.Read(), a, y.Read(), z.Size())); }

class Operator
{
protected:
   int64_t height = 0, width = 0;
public:
   Operator(int64_t s = 0) : height(s), width(s) { }
   virtual ~Operator() { }
   int64_t Height() const { return height; }
   virtual void Mult(const Vector &x, Vector &y) const = 0;
   virtual void MultTranspose(const Vector &, Vector &) const { LAGHOS_ABORT("Operator::MultTranspose() is not overloaded!"); }
};

class TimeDependentOperator : public Operator
{
protected:
   double t = 0.0;
public:
   TimeDependentOperator(int64_t n) : Operator(n) { }
   void SetTime(double t_) { t = t_; }
};

namespace hydrodynamics {

// reference laghos_assembly.hpp:31-62 — the arrays live in the lagb_ctx.
struct QuadratureData
{
   lagb_ctx *ctx = nullptr;
   double h0 = 0.0;
   double dt_est = std::numeric_limits<double>::infinity(); // host mirror after GetTimeStepEstimate
   double *stressJinvT() const { return lagb_qdata_ptr(ctx, 0); }
   double *rho0DetJ0w() const { return lagb_qdata_ptr(ctx, 1); }
   double *Jac0inv() const { return lagb_qdata_ptr(ctx, 2); }
};

struct TimingData { lagb_timing t; }; // reference laghos_solver.hpp:39-56

// reference laghos_assembly.hpp:115-131
class MassPAOperator : public Operator
{
   lagb_ctx *ctx;
   bool is_l2;
   mutable int ess_comp = -1;   // SetEssentialTrueDofs(c_tdofs[c]) selects the component's list
public:
   MassPAOperator(lagb_ctx *c, bool l2, int64_t size) : Operator(size), ctx(c), is_l2(l2) { }
   virtual void Mult(const Vector &x, Vector &y) const;
   void MultFull(const Vector &x, Vector &y) const;
   void SetEssentialTrueDofs(int comp) { ess_comp = comp; }
   void EliminateRHS(Vector &) const { /* folded into the PCG: see pcg.cuh */ }
   int EssComp() const { return ess_comp; }
   bool IsL2() const { return is_l2; }
   lagb_ctx *Ctx() const { return ctx; }
};

// reference laghos_assembly.hpp:94-112
class ForcePAOperator : public Operator
{
   lagb_ctx *ctx;
public:
   ForcePAOperator(lagb_ctx *c) : Operator(), ctx(c) { }
   virtual void Mult(const Vector &x, Vector &y) const { LAGHOS_CHECK(lagb_force_mult(ctx, x.Read(), y.Write())); }
   virtual void MultTranspose(const Vector &x, Vector &y) const { LAGHOS_CHECK(lagb_force_mult_transpose(ctx, x.Read(), y.Write())); }
};

// reference laghos_solver.hpp:58-93
class QUpdate
{
   lagb_ctx *ctx;
   double cfl;
public:
   QUpdate(lagb_ctx *c, double cfl_) : ctx(c), cfl(cfl_) { }
   void UpdateQuadratureData(const Vector &S, QuadratureData &) { LAGHOS_CHECK(lagb_qupdate_async(ctx, S.Read(), cfl)); }
};

} // namespace hydrodynamics

class Solver : public Operator
{
public:
   bool iterative_mode = false;
   Solver(int64_t s = 0) : Operator(s) { }
   virtual void SetOperator(const Operator &op) = 0;
};

// MFEM CGSolver surface used by the reference (laghos_solver.cpp:270-283, 388-392)
class CGSolver : public Solver
{
   const hydrodynamics::MassPAOperator *oper = nullptr;
   double rel_tol = 0.0, abs_tol = 0.0;
   int max_iter = 10, print_level = -1;
   mutable int final_iter = 0;
   bool have_prec = false;
public:
   CGSolver() { iterative_mode = true; }
   virtual void SetOperator(const Operator &op);
   void SetPreconditionerJacobi() { have_prec = true; } // OperatorJacobiSmoother(VMassPA->GetBF(), empty)
   void SetRelTol(double r) { rel_tol = r; }
   void SetAbsTol(double a) { abs_tol = a; }
   void SetMaxIter(int m) { max_iter = m; }
   void SetPrintLevel(int p) { print_level = p; }
   int GetNumIterations() const { return final_iter; }
   virtual void Mult(const Vector &b, Vector &x) const;
};

namespace hydrodynamics {

// reference laghos_solver.hpp:97-205 / laghos_solver.cpp:104-540
class LagrangianHydroOperator : public TimeDependentOperator
{
protected:
   lagb_ctx *ctx;
   lagb_problem_info info;
   int dim, source_type;
   int64_t H1Vsize, L2Vsize;
   double cfl, cg_rel_tol;
   int cg_max_iter;
   bool batched_pcg;
   mutable QuadratureData qdata;
   mutable bool qdata_is_current = false;
   mutable bool state_pending = false;   // a background H2D copy of S is in flight (bench e2e): wait before the first read
   ForcePAOperator *ForcePA;
   MassPAOperator *VMassPA, *EMassPA;
   mutable CGSolver CG_VMass, CG_EMass;
   mutable QUpdate qupdate;
   mutable Vector one, rhs, e_rhs, e_source, accel_b;
public:
   LagrangianHydroOperator(lagb_ctx *ctx, const lagb_problem_info &info, double cfl, double cgt, int cgiter, bool batched);
   ~LagrangianHydroOperator();
   virtual void Mult(const Vector &S, Vector &dS_dt) const;
   void SolveVelocity(const Vector &S, Vector &dS_dt) const;
   void SolveEnergy(const Vector &S, const Vector &v, Vector &dS_dt) const;
   void UpdateMesh(const Vector &) const { }
   double GetTimeStepEstimate(const Vector &S) const;
   void ResetTimeStepEstimate() const;
   void ResetQuadratureData() const { qdata_is_current = false; }
   // reference laghos_solver.cpp:639-697 (the MPI_Allreduce is inside the C-ABI call)
   double InternalEnergy(const Vector &e) const { double v; LAGHOS_CHECK(lagb_internal_energy(ctx, e.Read(), &v)); return v; }
   double KineticEnergy(const Vector &v_) const { double v; LAGHOS_CHECK(lagb_kinetic_energy(ctx, v_.Read(), &v)); return v; }
   // reference laghos_solver.cpp:542-563 (x: the mesh nodes the density is evaluated on; rho: L2 vector)
   void ComputeDensity(const Vector &x, Vector &rho) const { LAGHOS_CHECK(lagb_compute_density(ctx, x.Read(), rho.Write())); }
   void StatePending() const { state_pending = true; }
   void WaitState() const { if (state_pending) { LAGHOS_CHECK(lagb_wait_copies(ctx)); state_pending = false; } }
   void UpdateQuadratureData(const Vector &S) const;
   int64_t GetH1VSize() const { return H1Vsize; }
   // reference PrintTimingData (laghos_solver.cpp:699-796): fom[0]=total, 1=CG(H1), 2=forces, 3=qdata, 4=T_total
   void GetFOM(long long steps_x_stages, double fom[5], lagb_timing &tm) const;
   void PrintTimingData(bool IamRoot, long long steps, bool fom) const;
};

} // namespace hydrodynamics

class ODESolver
{
protected:
   TimeDependentOperator *f = nullptr;
public:
   virtual ~ODESolver() { }
   virtual void Init(TimeDependentOperator &f_) { f = &f_; }
   virtual void Step(Vector &x, double &t, double &dt) = 0;
};
class ForwardEulerSolver : public ODESolver { Vector dxdt; public: void Init(TimeDependentOperator &f_) override; void Step(Vector &x, double &t, double &dt) override; };
class RK2Solver : public ODESolver { double a; Vector dxdt, x1; public: RK2Solver(double a_ = 2./3.) : a(a_) { } void Init(TimeDependentOperator &f_) override; void Step(Vector &x, double &t, double &dt) override; };
class RK3SSPSolver : public ODESolver { Vector y, k; public: void Init(TimeDependentOperator &f_) override; void Step(Vector &x, double &t, double &dt) override; };
class RK4Solver : public ODESolver { Vector y, k, z; public: void Init(TimeDependentOperator &f_) override; void Step(Vector &x, double &t, double &dt) override; };
// MFEM ExplicitRKSolver with the RK6Solver tables (8 stages, 6th order; reference laghos.cpp:529, -s 6)
class RK6Solver : public ODESolver { Vector y, k[8]; public: void Init(TimeDependentOperator &f_) override; void Step(Vector &x, double &t, double &dt) override; };
// reference laghos_solver.hpp:232-255, laghos_solver.cpp:1429-1487
class HydroODESolver : public ODESolver
{
protected:
   hydrodynamics::LagrangianHydroOperator *hydro_oper = nullptr;
public:
   void Init(TimeDependentOperator &f_) override;
};
class RK2AvgSolver : public HydroODESolver { Vector V, dS_dt, S0; public: void Init(TimeDependentOperator &f_) override; void Step(Vector &S, double &t, double &dt) override; };

} // namespace laghos

// ---------------------------------------------------------------------------
// C entry point: the reference driver's run (laghos.cpp main) for rectilinear meshes.
// ---------------------------------------------------------------------------
extern "C" {
typedef struct lagb_run_options
{
   const char *mesh;        // data/ mesh stem
   int rs, problem, ok, ot, oq;
   double blast_scale;      // E0/2^dim (parallel driver) or 0.25 (serial driver)
   int impose_visc;
   int ode_solver_type;     // -s
   double t_final;          // -tf
   int max_tsteps;          // -ms
   double cfl, cg_tol;      // -cfl -cgt
   int cg_max_iter;         // -cgm
   int batched_pcg;         // 1: all velocity components in one batched PCG (default), 0: sequential like the reference
   int kernel_variant;      // lagb_ctx_desc.kernel_variant
   int device;
   int verbose, vis_steps;
   int e2e_host_state;      // 1: the state S lives in pinned host memory and is copied H2D before / D2H after every step
   int warmup_steps;        // steps before the timers are reset (bench)
   // multi-rank (element boxes of the global mesh); nranks <= 1: single GPU
   int rank, nranks;
   int pgrid[3];
   const unsigned char *nccl_id; // 128 bytes from lagb_nccl_unique_id (rank 0), broadcast by the launcher
   int profile_mass;        // 1: CUDA-event timing of every H1 mass-apply launch (bench roofline)
   // output files at every vis_steps-th accepted step and at the last one (laghos.cpp:845-900), off by default
   int gfprint;             // -print: <basename>_<ti>_{mesh,rho,v,e} (per-rank suffix .<rank:06d> when nranks > 1)
   int visit;               // -visit: <basename>_<ti:06d>.mfem_root + <basename>_<ti:06d>/<field>.<rank:06d>
   const char *basename;    // -k, default "results/Laghos" (the directory must exist)
   int check_exact_sedov;   // -err: problem 1: L2 error of the density against the exact Sedov solution at t_final (laghos.cpp:1009-1085)
   int v_error;             // 1: problems 0 / 4: L_inf, L_1, L_2 velocity errors at the end of the run (laghos.cpp:970-982; host)
   int check;               // --checks (laghos.cpp:904-926): |e| against the reference's table after every accepted step, two hits required
   double check_eps;        // relative tolerance of the check; <= 0: the reference's 1e-13
} lagb_run_options;

typedef struct lagb_run_result
{
   int steps, ti_last, stages;
   double t, dt, e_norm;
   double fom[5];           // total, CG(H1), forces, qdata, T_major (seconds)
   lagb_timing timing;
   double wall_seconds;     // host wall clock of the timed steps
   double device_seconds;   // CUDA-event time of the timed steps (whole loop)
   int64_t h2d_bytes_per_step, d2h_bytes_per_step;
   int64_t kernel_launches;
   int n_hist;
   int64_t ndofs_h1_global, ndofs_l2_global, ne_global;
   double mass_kernel_seconds;      // profile_mass: summed duration and count of the H1 mass-apply launches
   int64_t mass_kernel_launches;
   int64_t mass_kernel_ncomp;       // components per launch (3 = batched PCG)
   double work_mdof;                // numerator of the FOM: 1e-6 * (H1 dofs x CG its + (H1+L2) dofs x stages + quad points x updates)
   double energy_init, energy_final; // IE + KE before / after the run (laghos.cpp:664-665, 956-962 "Energy diff")
   double v_err[3];                 // v_error: L_inf, L_1, L_2 of v - v0(x) (problems 0 and 4), else 0
   double density_l2_err;           // check_exact_sedov: "Density L2 error", else 0
   int checks;                      // check: number of table entries that fired (2 for a complete run)
} lagb_run_result;

// hist: [2*hist_cap] (ti, |e|) pairs after every accepted step; S_out (optional): final state on the host
int lagb_laghos_run(const lagb_run_options *opt, lagb_run_result *res, double *hist, int hist_cap, double *S_out);
void lagb_run_options_default(lagb_run_options *opt);
}
