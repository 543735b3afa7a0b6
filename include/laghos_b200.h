/*
 * laghos_b200 — C ABI of the B200-native partial-assembly hot path of Laghos.
 *
 * The reference has no FFI layer: its hot path sits behind MFEM's virtual
 * mfem::Operator / mfem::Solver interfaces.  Every entry
 * point below names the reference interface it replaces (file:line in
 * /root/reference); the C++ classes in laghos_b200/shim/ keep the reference's class
 * and method names and forward to these calls.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures; `stream` is a cudaStream_t
 *     passed as void* (NULL = default stream);
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers;
 *   - all device work is enqueued asynchronously on the context's stream unless the
 *     function returns a host scalar (documented per call);
 *   - return value: 0 = ok, non-zero = error; lagb_last_error() gives the message.
 *     The shim turns a non-zero status into the reference's MFEM_ABORT behaviour.
 *   - layouts are the reference's: E-vector index
 *     ix + D1D*(iy + D1D*iz); L-vectors Ordering::byNODES; quadrature index
 *     q = qx + Q1D*(qy + Q1D*qz); stressJinvT[(e*NQ+q) + NE*NQ*(g + dim*c)];
 *     Jac0inv[i + dim*(j + dim*(e*NQ+q))]; state S = (x | v | e).
 */
#ifndef LAGHOS_B200_H
#define LAGHOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAGB_OK 0
#define LAGB_ERR_INVALID 1   /* bad argument / unknown kernel id (reference: MFEM_ABORT("Unknown kernel"), laghos_assembly.cpp:549-553) */
#define LAGB_ERR_CUDA 2
#define LAGB_ERR_NCCL 3
#define LAGB_ERR_STATE 4

const char *lagb_last_error(void);
/* number of CUDA kernels launched by this library since process start (bench.py `gpu_launches`) */
int64_t lagb_kernel_launch_count(void);

/* ------------------------------------------------------------------------- */
/* Host-side problem setup (mesh, tables, initial conditions).  Not timed.    */
/* Replaces reference laghos.cpp:380-656 for rectilinear meshes.              */
/* ------------------------------------------------------------------------- */
typedef struct lagb_problem lagb_problem;

typedef struct lagb_problem_info
{
   int32_t dim, NE, D1D, L1D, Q1D, ND, NL, NQ;
   int32_t nelem[3];       /* elements per axis */
   int32_t n1[3];          /* H1 lattice extents */
   int64_t ndofs_h1;       /* scalar H1 dofs; H1 vector size = dim*ndofs_h1 */
   int64_t ndofs_l2;
   int32_t use_visc, use_vort, source;
   int32_t ness[3];        /* essential scalar dofs per velocity component */
} lagb_problem_info;

/* mesh_name: stem of a reference data/ mesh (cube01_hex, square01_quad, box01_hex,
 * rectangle01_quad, square_gresho, rt2D).  blast_scale is the value handed to the
 * Sedov DeltaCoefficient (E0/2^dim in laghos.cpp:603-604; 0.25 in serial/laghos.cpp:101). */
int lagb_problem_create(lagb_problem **out, const char *mesh_name, int rs, int problem,
                        int ok, int ot, int oq, double blast_scale, int impose_visc);
/* generic rectilinear mesh: breakpoints per axis before the rs uniform refinements */
int lagb_problem_create_rect(lagb_problem **out, int dim,
                             const double *bx, int nbx, const double *by, int nby,
                             const double *bz, int nbz, int rs, int problem,
                             int ok, int ot, int oq, double blast_scale, int impose_visc);
/* the same from a mesh FILE in the reference's format (MFEM mesh v1.0, reference data/<name>.mesh;
 * `-m <file>`, laghos.cpp:380-393): rectilinear quad / hex meshes with the reference's boundary
 * attribute convention (attribute k = faces of constant x_{k-1}); anything else is rejected. */
int lagb_problem_create_file(lagb_problem **out, const char *path, int rs, int problem,
                             int ok, int ot, int oq, double blast_scale, int impose_visc);
/* per-axis breakpoints of the (refined, global) mesh */
int lagb_problem_mesh_breaks(const lagb_problem *p, int axis, const double **brk, int32_t *n);
/* one rank's part of an element-partitioned run: the element box of
 * `rank` in the process grid pgrid[3]; boundary conditions and the Sedov delta refer
 * to the global mesh.  Replaces ParMesh(MPI_COMM_WORLD, mesh, partitioning)
 * (laghos.cpp:481) for Cartesian partitions. */
int lagb_problem_create_part(lagb_problem **out, const char *mesh_name, int rs, int problem,
                             int ok, int ot, int oq, double blast_scale, int impose_visc,
                             int rank, const int32_t pgrid[3]);
/* neighbour k of a partitioned problem: rank, exchange phase, shared scalar dofs */
int lagb_problem_nnbr(const lagb_problem *p);
int lagb_problem_nbr(const lagb_problem *p, int k, int32_t *rank, int32_t *phase, int32_t *n,
                     const int32_t **dofs);
const uint8_t *lagb_problem_owner_mask(const lagb_problem *p);   /* [ndofs_h1]; NULL if not partitioned */
void lagb_problem_destroy(lagb_problem *p);
int lagb_problem_get_info(const lagb_problem *p, lagb_problem_info *info);
/* read-only views into the problem's host arrays */
const int32_t *lagb_problem_h1_map(const lagb_problem *p);        /* [NE*ND] */
const int32_t *lagb_problem_ess(const lagb_problem *p, int c);    /* [ness[c]] */
const double *lagb_problem_S0(const lagb_problem *p);             /* [2*dim*ndofs_h1 + ndofs_l2] */
const double *lagb_problem_rho0_gf(const lagb_problem *p);        /* [ndofs_l2] Bernstein coefficients */
const double *lagb_problem_rho0_q(const lagb_problem *p);         /* [NE*NQ] analytic rho0 at quad points */
const double *lagb_problem_gamma(const lagb_problem *p);          /* [NE] */
const double *lagb_problem_qweights(const lagb_problem *p);       /* [NQ] */
/* 1D tables: which = 0:B 1:G (H1, [q + Q1D*d]), 2:BL (L2 Bernstein, [q + Q1D*l]), 3:qx 4:qw */
const double *lagb_problem_table(const lagb_problem *p, int which);

/* The reference driver's self-test `--checks` (laghos.cpp:904-926, 1403-1474; table it_norms :1441-1463).
 *   lagb_checks_entry : entry k (0, 1) of the table for (dim, problem): step index and |e|
 *   lagb_checks_step  : the reference's Checks(ti, |e|, chk): *chk is incremented when the table has an entry at this
 *                       step; returns LAGB_OK (no entry, or the entry matched) or LAGB_ERR_INVALID on a mismatch
 *                       (lagb_last_error() = "P<problem>, #<step>"); eps is the reference's 1e-13, both relative
 *                       errors must be below it
 * The run option `check` of lagb_laghos_run applies it after every accepted step and requires two hits at the end. */
int lagb_checks_entry(int dim, int problem, int k, int32_t *it, double *norm);
int lagb_checks_step(int dim, int problem, int ti, double e_norm, double eps, int32_t *chk);

/* End-of-run velocity error norms of the reference driver for problems 0 and 4 (laghos.cpp:970-982:
 * ComputeMaxError / ComputeL1Error / ComputeL2Error against the initial velocity field, which is the exact solution
 * there), from the HOST copy of the state S = (x | v | e): Gauss-Legendre rule of order 2 ok + 3 per element, exact
 * field evaluated at the current position of each point.  out = { L_inf, L_1, L_2, sum of squares behind L_2 };
 * a partitioned run reduces out[0] with max, out[1] and out[3] with sum, and takes L_2 = sqrt(out[3]). */
int lagb_problem_velocity_error(const lagb_problem *p, const double *h_S, double out[4]);

/* `-err` of the reference driver (laghos.cpp:1009-1085, problem 1 on the built-in mesh): the exact Taylor - von Neumann
 * - Sedov solution (own restatement of the published similarity solution, host/sedov_exact.hpp; the reference uses
 * sedov/sedov_sol.cpp) and the L2 error of the density field against it.
 *   lagb_sedov_exact_eval: rho, v, p at n radii at time t; alpha_override > 0 replaces the energy integral;
 *                          info = { alpha, r2, U, rho2, v2, p2 }
 *   lagb_problem_sedov_density_error: h_S = host state (positions used), h_rho = ComputeDensity field [ndofs_l2],
 *                          gamma / rho0 / blast_energy as the driver passes them (1.4, 1, E0);
 *                          out = { L2 error, sum of squares } (partitioned runs add out[1] and take the root) */
int lagb_sedov_exact_eval(int dim, double gamma, double rho0, double blast_energy, double t, double alpha_override,
                          int n, const double *r, double *rho, double *v, double *p, double info[6]);
int lagb_problem_sedov_density_error(const lagb_problem *p, const double *h_S, const double *h_rho, double t,
                                     double gamma, double rho0, double blast_energy, double out[2]);

/* Output files (host code, off the timed path): the reference's `-print` files
 * <basename>_<ti>_mesh / _rho / _v / _e (laghos.cpp:873-900) and its VisIt data collection
 * (laghos.cpp:866-871) in MFEM's text formats (mesh v1.0 with a `nodes` grid function, GridFunction::Save).
 * H1 fields (mesh nodes, velocity) are written element-wise in L2_T1_<dim>D_P<ok> (discontinuous
 * Gauss-Lobatto: exact for the path's H1 basis, independent of a reader's edge / face numbering), L2 fields
 * (e, rho) in the path's own L2_T2_<dim>D_P<ot> layout.  All pointers are HOST pointers.
 *   lagb_problem_write_mesh : h_x = H1 positions [dim*ndofs_h1] (NULL: the initial mesh)
 *   lagb_problem_write_field: kind 0 = H1 field with vdim components [vdim*ndofs_h1], 1 = L2 scalar [ndofs_l2]
 *   lagb_problem_write_print: the four `-print` files of step ti from the state S = (x | v | e) and rho
 *   lagb_problem_write_visit: <collection>_<cycle:06d>/{mesh,Density,Velocity,Specific Internal Energy}.<rank:06d>
 *                             and, on rank 0, <collection>_<cycle:06d>.mfem_root (h_rho may be NULL: no Density)
 * A partitioned problem writes its own element block (a valid serial mesh of that block). */
int lagb_problem_write_mesh(const lagb_problem *p, const double *h_x, const char *path, int precision);
int lagb_problem_write_field(const lagb_problem *p, int kind, int vdim, const double *h_f, const char *path,
                             int precision);
int lagb_problem_write_print(const lagb_problem *p, const char *basename, int ti, const double *h_S,
                             const double *h_rho, int precision);
int lagb_problem_write_visit(const lagb_problem *p, const char *collection, int cycle, double time, double time_step,
                             int rank, int nranks, const double *h_S, const double *h_rho, int precision);

/* ------------------------------------------------------------------------- */
/* Device context: the operators' shared state (QuadratureData and work        */
/* vectors).  Replaces the members of LagrangianHydroOperator that the PA      */
/* operators reference (laghos_solver.hpp:97-205, laghos_assembly.hpp:31-62).  */
/* ------------------------------------------------------------------------- */
typedef struct lagb_ctx lagb_ctx;

typedef struct lagb_ctx_desc
{
   int32_t dim, NE, D1D, L1D, Q1D;
   int64_t ndofs_h1;                 /* scalar H1 dofs on this rank (incl. shared) */
   const int32_t *h_h1_map;          /* [NE*D1D^dim] ElementRestriction, lexicographic */
   const int32_t *h_ess[3];          /* essential scalar dofs per component */
   int32_t ness[3];
   const double *h_B, *h_G, *h_BL;   /* 1D tables, [q + Q1D*d] */
   const double *h_qweights;         /* [NQ] */
   const double *h_gamma;            /* [NE] */
   int32_t use_visc, use_vort;
   int32_t device;                   /* CUDA device ordinal */
   int32_t kernel_variant;           /* 0 = tuned kernels where available, 1 = generic one-thread-per-element kernels */
   int32_t elem_grid[3];             /* optional hint: elements are numbered lexicographically (x fastest) on an
                                        nx*ny*nz grid; {0,0,0} = unstructured.  Only the batching of the mass apply
                                        (bricks vs consecutive elements) depends on it, never the result's value set */
} lagb_ctx_desc;

int lagb_ctx_create(lagb_ctx **out, const lagb_ctx_desc *desc, void *stream);
void lagb_ctx_destroy(lagb_ctx *ctx);
int lagb_ctx_sync(lagb_ctx *ctx);   /* cudaStreamSynchronize on the context stream */

/* t = 0 setup — reference Rho0DetJ0Vol (laghos_solver.cpp:1170-1261), h0
 * (:251-262), MassIntegrator(rho0).AssemblePA and the Jacobi diagonal
 * (laghos_assembly.cpp:92-95, laghos_solver.cpp:268-270).
 * d_x0: initial mesh nodes (H1 vector); d_rho0_gf: L2 Bernstein density;
 * d_rho0_q: coefficient values at quad points for the mass operator (NULL: use
 * the interpolated rho0_gf).  ne_global/vol_scale: for multi-rank h0 pass the
 * global element count (0 = local NE).  Synchronous (returns h0 on the host). */
int lagb_setup_qdata0(lagb_ctx *ctx, const double *d_x0, const double *d_rho0_gf,
                      const double *d_rho0_q, int64_t ne_global, double *h0_out);

/* MassPAOperator::Mult (laghos_assembly.cpp:117-121): y = M x on the scalar H1
 * space, then y[ess(comp)] = 0.  comp = -1: MultFull (laghos_assembly.hpp:128). */
int lagb_vmass_mult(lagb_ctx *ctx, int comp, const double *d_x, double *d_y);
/* the same for all `dim` components of an H1 vector at once (byNODES), no essential-dof
 * zeroing: y_c = M x_c.  One pass over the quadrature data for all components (the kernel
 * the batched PCG runs every iteration). */
int lagb_vmass_mult_all(lagb_ctx *ctx, const double *d_x, double *d_y);
/* OperatorJacobiSmoother diagonal (laghos_solver.cpp:268-270): d_diag[ndofs_h1] */
int lagb_vmass_diag(lagb_ctx *ctx, double *d_diag);
/* MassPAOperator(L2)::Mult (laghos_solver.cpp:179): block-diagonal Bernstein mass */
int lagb_emass_mult(lagb_ctx *ctx, const double *d_x, double *d_y);

/* ForcePAOperator::Mult (laghos_assembly.cpp:557-565): d_e L2 vector -> d_v H1 vector (byNODES) */
int lagb_force_mult(lagb_ctx *ctx, const double *d_e, double *d_v);
/* ForcePAOperator::MultTranspose (laghos_assembly.cpp:965-973) */
int lagb_force_mult_transpose(lagb_ctx *ctx, const double *d_v, double *d_e);

/* QUpdate::UpdateQuadratureData (laghos_solver.cpp:1354-1411): recomputes
 * stressJinvT from the state S and returns min(dt_est_in, min_q dt_q) on the host
 * (the reference's q_dt_est.Min(), a device->host sync, :1406). */
int lagb_qupdate(lagb_ctx *ctx, const double *d_S, double cfl, double dt_est_in, double *h_dt_est_out);
/* The same without a host round trip per call: the running minimum
 * (QuadratureData::dt_est, laghos_assembly.hpp:55) stays in a device scalar.
 * lagb_dt_est_set = ResetTimeStepEstimate (laghos_solver.cpp:537-540, v = +inf);
 * lagb_qupdate_async mins into it; lagb_dt_est_read = GetTimeStepEstimate's
 * read-back including the MPI_Allreduce(MIN) of laghos_solver.cpp:533. */
int lagb_dt_est_set(lagb_ctx *ctx, double v);
int lagb_qupdate_async(lagb_ctx *ctx, const double *d_S, double cfl);
int lagb_dt_est_read(lagb_ctx *ctx, double *h_dt_est_out);

/* CG_VMass.Mult(B, X) (laghos_solver.cpp:388; MFEM CGSolver + OperatorJacobiSmoother,
 * iterative_mode = true): Jacobi-PCG on the scalar velocity mass matrix with the
 * essential dofs of `comp` eliminated.  d_b is used as given except that entries at
 * essential dofs are treated as zero (EliminateRHS, laghos_assembly.cpp:112-115).
 * d_x: initial guess in, solution out.  h_iters: MFEM's GetNumIterations(). */
int lagb_pcg_vmass(lagb_ctx *ctx, int comp, const double *d_b, double *d_x,
                   double rel_tol, int max_iter, int *h_iters);
/* all `dim` component solves of SolveVelocity (laghos_solver.cpp:363-398) in one
 * batched PCG: d_rhs and d_dv are H1 vectors (byNODES); the quadrature data D is
 * read once per iteration for all components.  h_iters[c] per component. */
int lagb_pcg_vmass_all(lagb_ctx *ctx, const double *d_rhs, double *d_dv,
                       double rel_tol, int max_iter, int *h_iters);
/* The same with a ZERO initial guess (MFEM CGSolver with iterative_mode = false: x = 0, r = b; the reference's
 * SolveVelocity starts every solve from dv = 0, laghos_solver.cpp:363-398 with dS_dt = 0): d_dv is output only, and
 * the operator application to the initial guess is skipped.  Same iterates as lagb_pcg_vmass_all on a zeroed d_dv. */
int lagb_pcg_vmass_all_x0(lagb_ctx *ctx, const double *d_rhs, double *d_dv,
                          double rel_tol, int max_iter, int *h_iters);
/* CG_EMass.Mult(e_rhs, de) (laghos_solver.cpp:481): unpreconditioned CG,
 * iterative_mode = false. */
int lagb_cg_emass(lagb_ctx *ctx, const double *d_b, double *d_x,
                  double rel_tol, int max_iter, int *h_iters);

/* LagrangianHydroOperator::InternalEnergy / KineticEnergy (laghos_solver.cpp:639-697):
 * sum_q rho0 detJ0 w e(q)  and  1/2 sum_q rho0 detJ0 w |v(q)|^2, summed over the ranks
 * (the reference's MPI_Allreduce, :663, :693).  Evaluated as 1^t (M_L2 e) and 1/2 v^t (M_H1 v)
 * with the PA mass kernels (partition of unity of the Bernstein basis; the mass coefficient equals
 * rho0DetJ0w for the element-wise constant densities of the reference's problems).  Synchronous. */
int lagb_internal_energy(lagb_ctx *ctx, const double *d_e, double *h_out);
int lagb_kinetic_energy(lagb_ctx *ctx, const double *d_v, double *h_out);

/* LagrangianHydroOperator::ComputeDensity (laghos_solver.cpp:542-563): the L2 density field on the current
 * mesh, per element rho_z = Mrho(x)^-1 rhs with the DensityIntegrator right-hand side
 * (laghos_assembly.cpp:26-41).  Diagnostics path (visualisation / -err), synchronous.
 * d_x: current mesh nodes (H1 vector); d_rho: [ndofs_l2]. */
int lagb_compute_density(lagb_ctx *ctx, const double *d_x, double *d_rho);

/* 2D Taylor-Green energy source (laghos_solver.cpp:455-465): d_esrc[ndofs_l2] */
int lagb_taylor_source(lagb_ctx *ctx, const double *d_x, double *d_esrc);

/* QuadratureData accessors (laghos_assembly.hpp:31-62): which = 0 stressJinvT,
 * 1 rho0DetJ0w, 2 Jac0inv, 3 mass coefficient D, 4 H1 mass diagonal */
double *lagb_qdata_ptr(lagb_ctx *ctx, int which);
double lagb_qdata_h0(const lagb_ctx *ctx);
int lagb_qdata_set_h0(lagb_ctx *ctx, double h0);

/* Device memory for the shim's Vector (MFEM's Memory<double> with device residency
 * made explicit).  Copies are enqueued on the context stream; lagb_memcpy_d2h
 * synchronises before returning, lagb_memcpy_h2d_async does not (h_src must stay
 * valid until lagb_ctx_sync; use pinned memory from lagb_host_alloc_pinned). */
int lagb_dev_malloc(lagb_ctx *ctx, double **d_out, int64_t n);
int lagb_dev_free(lagb_ctx *ctx, double *d_ptr);
int lagb_memcpy_h2d(lagb_ctx *ctx, double *d_dst, const double *h_src, int64_t n);        /* synchronous */
int lagb_memcpy_h2d_async(lagb_ctx *ctx, double *d_dst, const double *h_src, int64_t n);
int lagb_memcpy_d2h(lagb_ctx *ctx, double *h_dst, const double *d_src, int64_t n);        /* synchronous */
/* Pipelined state transfers (bench e2e): copies run on the context's COPY stream and overlap the
 * compute stream.  lagb_memcpy_h2d_bg: the copy starts once the compute work enqueued so far has
 * finished (it overwrites d_dst) and after earlier background copies; kernels that read d_dst must be
 * preceded by lagb_wait_copies.  lagb_memcpy_d2h_bg: the copy starts once the compute work enqueued
 * so far has finished; h_dst is valid after lagb_ctx_sync (which also drains the copy stream). */
int lagb_memcpy_h2d_bg(lagb_ctx *ctx, double *d_dst, const double *h_src, int64_t n);
int lagb_memcpy_d2h_bg(lagb_ctx *ctx, double *h_dst, const double *d_src, int64_t n);
int lagb_wait_copies(lagb_ctx *ctx);   /* the compute stream waits for the background copies enqueued so far */
int lagb_host_alloc_pinned(double **h_out, int64_t n);
int lagb_host_free_pinned(double *h_ptr);

/* Small vector kernels used by the shim's Vector class and ODE solvers
 * (MFEM Vector::operator=, Add, add(), Neg, operator*): all on the ctx stream. */
int lagb_vec_fill(lagb_ctx *ctx, double *d_y, double a, int64_t n);
int lagb_vec_copy(lagb_ctx *ctx, double *d_y, const double *d_x, int64_t n);
int lagb_vec_axpby(lagb_ctx *ctx, double *d_z, double a, const double *d_x, double b, const double *d_y, int64_t n); /* z = a x + b y */
int lagb_vec_dot(lagb_ctx *ctx, const double *d_x, const double *d_y, int64_t n, double *h_out); /* synchronous */

/* ------------------------------------------------------------------------- */
/* Multi-GPU: element-partitioned ranks, one context per GPU.   */
/* ------------------------------------------------------------------------- */
/* NCCL bootstrap: rank 0 calls lagb_nccl_unique_id, the 128 bytes are broadcast by
 * the launcher (torch.distributed), every rank calls lagb_ctx_comm_init. */
int lagb_nccl_unique_id(uint8_t id_out[128]);
/* shared-dof description, one entry per neighbour rank in a fixed exchange order:
 * nbr_rank[k], and for each k the list of local scalar dof ids shared with that
 * rank (same order on both sides).  exchange_phase[k] groups exchanges that may
 * run concurrently (Cartesian partitions: phase = axis, so that edge and corner
 * dofs are summed by three successive face exchanges).  h_owner_mask[i] = 1 if
 * this rank owns scalar dof i in inner products (each shared dof is owned by
 * exactly one rank). */
int lagb_ctx_comm_init(lagb_ctx *ctx, const uint8_t id[128], int rank, int nranks,
                       int nnbr, const int32_t *nbr_rank, const int32_t *exchange_phase,
                       const int32_t *nshared, const int32_t *const *h_shared_dofs,
                       const uint8_t *h_owner_mask);

/* all-reduce of a few host scalars over the context's ranks (MPI_Allreduce call
 * sites of SURVEY.md 2.3: energies, |e|^2, timers): op 0 = sum, 1 = min, 2 = max.
 * No-op for a single rank.  Synchronous. */
int lagb_allreduce_host(lagb_ctx *ctx, double *h_vals, int n, int op);

/* ------------------------------------------------------------------------- */
/* Timers and counters of the reference's TimingData (laghos_solver.hpp:39-56)  */
/* ------------------------------------------------------------------------- */
typedef struct lagb_timing
{
   double t_cgH1, t_cgL2, t_force, t_qdata;   /* seconds, device-timed (CUDA events) */
   int64_t H1iter, L2iter, quad_tstep;
} lagb_timing;
int lagb_timing_get(lagb_ctx *ctx, lagb_timing *out);
int lagb_timing_reset(lagb_ctx *ctx);
/* Measurement aids (bench.py): a CUDA-event stopwatch on the context stream, and per-launch
 * CUDA-event timing of the H1 mass-apply kernel inside the PCG (the dominant kernel; its
 * average duration is the denominator of the roofline figure).  lagb_timing_reset clears both. */
int lagb_stopwatch_start(lagb_ctx *ctx);
int lagb_stopwatch_stop(lagb_ctx *ctx, double *seconds);   /* synchronises */
int lagb_profile_mass(lagb_ctx *ctx, int enable);
int lagb_profile_mass_get(lagb_ctx *ctx, double *seconds, int64_t *launches);
/* kernel tuning knobs (tools/microbench.py): key 0 = launch variant of the legacy 3-component mass apply,
 * 1 = Force/Force^T, 2 = QUpdate, 3 = legacy 1-component mass, 4 = brick mass apply variant,
 * 5 = 1: no programmatic dependent launch between the colours, 6 = mass-apply path (0 / 1: direct gather with
 * atomic scatter, the default; 2 / 3: coloured brick kernels with a fixed summation order, measured slower),
 * 10 = 1: the reference's global CG for the energy solve, 11 = 1: NCCL instead of peer-memory exchanges */
int lagb_tune_set(lagb_ctx *ctx, int key, int value);
/* HOST only (no CUDA call): builds the coloured brick schedule of the mass apply for the gather map
 * h_map [NE*ND] (grid = structured element grid hint or NULL, NB = elements per batch) and verifies
 * the invariants the kernels rely on (every element once, colours conflict-free, exactly one first
 * writer per dof, CSR consistent).  stats: nbatch, ncolors, ntables, max unique, padded unique,
 * touched dofs, brick shape bx+100*by+10000*bz, total unique entries.  Replaces nothing in the
 * reference: MFEM's ElementRestriction builds its offsets/indices arrays at the same point
 * (laghos_assembly.cpp:133-134). */
int lagb_host_batch_plan_check(const int32_t *h_map, int NE, int ND, int64_t ndofs, const int32_t grid[3],
                               int NB, int64_t stats[8]);

#ifdef __cplusplus
}
#endif
#endif /* LAGHOS_B200_H */
