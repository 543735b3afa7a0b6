"""Python binding of the C ABI (include/laghos_b200.h) for tests, bench and the launcher.

Names mirror the reference's operators: ``Context.vmass_mult`` is
``MassPAOperator::Mult`` (reference laghos_assembly.cpp:117-121), ``force_mult`` /
``force_mult_transpose`` are ``ForcePAOperator::Mult/MultTranspose`` (:557-565, :965-973),
``qupdate`` is ``QUpdate::UpdateQuadratureData`` (laghos_solver.cpp:1354-1411),
``pcg_vmass`` is ``CG_VMass.Mult`` (:388) and ``run`` is the driver loop (laghos.cpp:706-778).
Device vectors are float64 CUDA torch tensors; torch is plumbing only.
"""
import ctypes as C
import numpy as np

from ._lib import load_library, ProblemInfo, CtxDesc, Timing, RunOptions, RunResult, c_double_p


class LagbError(RuntimeError):
    pass


def _check(lib, rc):
    if rc != 0:
        raise LagbError(lib.lagb_last_error().decode())


def sedov_exact(dim, t, r, gamma=1.4, rho0=1.0, blast_energy=1.0, alpha=0.0):
    """exact Sedov blast wave at radii r: (rho, v, p, info = [alpha, r2, U, rho2, v2, p2]); host/sedov_exact.hpp"""
    lib = load_library()
    r = np.ascontiguousarray(r, dtype=np.float64)
    rho, v, p, info = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r), np.zeros(6)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    _check(lib, lib.lagb_sedov_exact_eval(dim, gamma, rho0, blast_energy, t, alpha, r.size, vp(r), vp(rho), vp(v), vp(p),
                                          vp(info)))
    return rho, v, p, info


def mesh_dim(mesh):
    """dimension of a named mesh (data/ stems, hexbox_PxQxR, the reference's built-in `default[_2d | _NXxNY[xNZ]...]`)"""
    if mesh in ("square01_quad", "rectangle01_quad", "square_gresho", "rt2D", "default_2d"):
        return 2
    if mesh.startswith("default_"):
        return len(mesh.split("_")[1].split("x"))
    return 3


class Problem:
    """Host-side setup (mesh, tables, initial conditions): reference laghos.cpp:380-656."""

    def __init__(self, mesh="cube01_hex", rs=0, problem=1, ok=2, ot=1, oq=-1, blast_scale=None, impose_visc=False,
                 rank=0, pgrid=None, mesh_file=None, dim=None):
        self.lib = load_library()
        if dim is None:
            dim = mesh_dim(mesh)
        if blast_scale is None:
            blast_scale = 1.0 / 2 ** dim  # E0 = 1 (laghos.cpp:166, 603-604)
        self.args = dict(mesh=mesh, rs=rs, problem=problem, ok=ok, ot=ot, oq=oq, blast_scale=blast_scale,
                         impose_visc=impose_visc)
        h = C.c_void_p()
        self.pgrid = pgrid
        if mesh_file is not None:
            # reference `-m <file>`: MFEM mesh v1.0, rectilinear (pass dim for the default blast scale)
            _check(self.lib, self.lib.lagb_problem_create_file(C.byref(h), str(mesh_file).encode(), rs, problem, ok, ot,
                                                               oq, blast_scale, int(impose_visc)))
        elif pgrid is None:
            _check(self.lib, self.lib.lagb_problem_create(C.byref(h), mesh.encode(), rs, problem, ok, ot, oq,
                                                          blast_scale, int(impose_visc)))
        else:
            pg = (C.c_int32 * 3)(*pgrid)
            _check(self.lib, self.lib.lagb_problem_create_part(C.byref(h), mesh.encode(), rs, problem, ok, ot, oq,
                                                               blast_scale, int(impose_visc), rank, C.byref(pg)))
        self.h = h
        info = ProblemInfo()
        _check(self.lib, self.lib.lagb_problem_get_info(h, C.byref(info)))
        self.info = info
        for k in ("dim", "NE", "D1D", "L1D", "Q1D", "ND", "NL", "NQ", "ndofs_h1", "ndofs_l2", "use_visc", "use_vort",
                  "source"):
            setattr(self, k, int(getattr(info, k)))
        self.h1_vsize = self.dim * self.ndofs_h1
        self.s_size = 2 * self.h1_vsize + self.ndofs_l2

    def _arr(self, ptr, n, dtype):
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

    @property
    def S0(self):
        return self._arr(self.lib.lagb_problem_S0(self.h), self.s_size, np.float64)

    @property
    def rho0_gf(self):
        return self._arr(self.lib.lagb_problem_rho0_gf(self.h), self.ndofs_l2, np.float64)

    @property
    def rho0_q(self):
        return self._arr(self.lib.lagb_problem_rho0_q(self.h), self.NE * self.NQ, np.float64)

    @property
    def gamma(self):
        return self._arr(self.lib.lagb_problem_gamma(self.h), self.NE, np.float64)

    @property
    def h1_map(self):
        return self._arr(self.lib.lagb_problem_h1_map(self.h), self.NE * self.ND, np.int32)

    def ess(self, c):
        n = int(self.info.ness[c])
        return self._arr(self.lib.lagb_problem_ess(self.h, c), n, np.int32) if n else np.zeros(0, np.int32)

    def neighbours(self):
        """[(rank, phase, shared scalar dof ids)] of an element-partitioned problem."""
        out = []
        for k in range(self.lib.lagb_problem_nnbr(self.h)):
            r, ph, n = C.c_int32(), C.c_int32(), C.c_int32()
            ptr = C.POINTER(C.c_int32)()
            _check(self.lib, self.lib.lagb_problem_nbr(self.h, k, C.byref(r), C.byref(ph), C.byref(n), C.byref(ptr)))
            out.append((r.value, ph.value, self._arr(ptr, n.value, np.int32)))
        return out

    @property
    def owner_mask(self):
        ptr = self.lib.lagb_problem_owner_mask(self.h)
        if not ptr:
            return np.ones(self.ndofs_h1, dtype=np.uint8)
        return self._arr(ptr, self.ndofs_h1, np.uint8)

    def mesh_breaks(self, axis):
        ptr = c_double_p()
        n = C.c_int32()
        _check(self.lib, self.lib.lagb_problem_mesh_breaks(self.h, axis, C.byref(ptr), C.byref(n)))
        return self._arr(ptr, n.value, np.float64)

    def table(self, which, n):
        return self._arr(self.lib.lagb_problem_table(self.h, which), n, np.float64)

    def velocity_error(self, S):
        """(L_inf, L_1, L_2) error of v against the initial (exact) velocity field: laghos.cpp:970-982, problems 0 / 4"""
        keep, pS = self._hp(S, self.s_size)
        out = (C.c_double * 4)()
        _check(self.lib, self.lib.lagb_problem_velocity_error(self.h, pS, out))
        return out[0], out[1], out[2]

    def sedov_density_error(self, S, rho, t, gamma=1.4, rho0=1.0, blast_energy=1.0):
        """`-err`: L2 error of the density field against the exact Sedov solution at time t (laghos.cpp:1009-1085)"""
        kS, pS = self._hp(S, self.s_size)
        kr, pr = self._hp(rho, self.ndofs_l2)
        out = (C.c_double * 2)()
        _check(self.lib, self.lib.lagb_problem_sedov_density_error(self.h, pS, pr, t, gamma, rho0, blast_energy, out))
        return out[0]

    # ---- output files (reference -print / -visit, laghos.cpp:866-900); host arrays ----
    @staticmethod
    def _hp(a, n):
        if a is None:
            return None, None
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.size != n:
            raise ValueError(f"expected {n} values, got {a.size}")
        return a, a.ctypes.data_as(C.c_void_p)

    def write_mesh(self, path, x=None, precision=8):
        keep, px = self._hp(x, self.dim * self.ndofs_h1)
        _check(self.lib, self.lib.lagb_problem_write_mesh(self.h, px, str(path).encode(), precision))

    def write_field(self, path, f, kind, vdim=1, precision=8):
        """kind 0: H1 field with vdim components [vdim*ndofs_h1]; kind 1: L2 scalar [ndofs_l2]"""
        keep, pf = self._hp(f, vdim * self.ndofs_h1 if kind == 0 else self.ndofs_l2)
        _check(self.lib, self.lib.lagb_problem_write_field(self.h, kind, vdim, pf, str(path).encode(), precision))

    def write_print(self, basename, ti, S, rho, precision=8):
        kS, pS = self._hp(S, self.s_size)
        kr, pr = self._hp(rho, self.ndofs_l2)
        _check(self.lib, self.lib.lagb_problem_write_print(self.h, str(basename).encode(), ti, pS, pr, precision))

    def write_visit(self, collection, cycle, time, time_step, S, rho=None, rank=0, nranks=1, precision=8):
        kS, pS = self._hp(S, self.s_size)
        kr, pr = self._hp(rho, self.ndofs_l2)
        _check(self.lib, self.lib.lagb_problem_write_visit(self.h, str(collection).encode(), cycle, time, time_step,
                                                           rank, nranks, pS, pr, precision))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.lagb_problem_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Context:
    """Device context: QuadratureData + operators (reference laghos_solver.hpp:97-205)."""

    def __init__(self, problem, device=0, variant=0, setup=True, grid_hint=True):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise LagbError("no CUDA device: laghos_b200 has no CPU fallback")
        self.lib = problem.lib
        self.P = problem
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        d = CtxDesc()
        d.dim, d.NE, d.D1D, d.L1D, d.Q1D = problem.dim, problem.NE, problem.D1D, problem.L1D, problem.Q1D
        d.ndofs_h1 = problem.ndofs_h1
        lib = self.lib
        d.h_h1_map = C.cast(lib.lagb_problem_h1_map(problem.h), C.c_void_p)
        for c in range(problem.dim):
            d.h_ess[c] = C.cast(lib.lagb_problem_ess(problem.h, c), C.c_void_p)
            d.ness[c] = problem.info.ness[c]
        d.h_B = C.cast(lib.lagb_problem_table(problem.h, 0), C.c_void_p)
        d.h_G = C.cast(lib.lagb_problem_table(problem.h, 1), C.c_void_p)
        d.h_BL = C.cast(lib.lagb_problem_table(problem.h, 2), C.c_void_p)
        d.h_qweights = C.cast(lib.lagb_problem_qweights(problem.h), C.c_void_p)
        d.h_gamma = C.cast(lib.lagb_problem_gamma(problem.h), C.c_void_p)
        d.use_visc, d.use_vort, d.device, d.kernel_variant = problem.use_visc, problem.use_vort, device, variant
        if grid_hint:
            for k in range(3):   # the problem's meshes are Cartesian with lexicographic element numbering
                d.elem_grid[k] = (int(problem.info.n1[k]) - 1) // (problem.D1D - 1) if k < problem.dim else 1
        h = C.c_void_p()
        # the context runs on torch's current stream so that torch.cuda.Event timing sees it
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(lib, lib.lagb_ctx_create(C.byref(h), C.byref(d), C.c_void_p(stream)))
        self.h = h
        self.h0 = None
        if setup:
            self.setup()

    # -- helpers --
    def dev(self, a):
        return self.torch.as_tensor(np.ascontiguousarray(a), dtype=self.torch.float64).to(self.device)

    def empty(self, n):
        return self.torch.empty(int(n), dtype=self.torch.float64, device=self.device)

    def zeros(self, n):
        return self.torch.zeros(int(n), dtype=self.torch.float64, device=self.device)

    @staticmethod
    def _p(t):
        assert t.is_cuda and t.dtype.is_floating_point and t.element_size() == 8 and t.is_contiguous()
        return C.c_void_p(t.data_ptr())

    def setup(self, x0=None, rho0_gf=None, rho0_q=None):
        P = self.P
        x0 = self.dev(P.S0[:P.h1_vsize]) if x0 is None else x0
        rho0_gf = self.dev(P.rho0_gf) if rho0_gf is None else rho0_gf
        rho0_q = self.dev(P.rho0_q) if rho0_q is None else rho0_q
        h0 = C.c_double()
        _check(self.lib, self.lib.lagb_setup_qdata0(self.h, self._p(x0), self._p(rho0_gf), self._p(rho0_q), 0,
                                                    C.byref(h0)))
        self.h0 = h0.value
        return self.h0

    def sync(self):
        _check(self.lib, self.lib.lagb_ctx_sync(self.h))

    # -- operators --
    def vmass_mult(self, x, comp=-1):
        y = self.empty(self.P.ndofs_h1)
        _check(self.lib, self.lib.lagb_vmass_mult(self.h, comp, self._p(x), self._p(y)))
        return y

    def vmass_mult_all(self, x, y=None):
        y = self.empty(self.P.h1_vsize) if y is None else y
        _check(self.lib, self.lib.lagb_vmass_mult_all(self.h, self._p(x), self._p(y)))
        return y

    def tune(self, key, value):
        _check(self.lib, self.lib.lagb_tune_set(self.h, key, value))

    def vmass_diag(self):
        y = self.empty(self.P.ndofs_h1)
        _check(self.lib, self.lib.lagb_vmass_diag(self.h, self._p(y)))
        return y

    def emass_mult(self, x):
        y = self.empty(self.P.ndofs_l2)
        _check(self.lib, self.lib.lagb_emass_mult(self.h, self._p(x), self._p(y)))
        return y

    def force_mult(self, e):
        v = self.empty(self.P.h1_vsize)
        _check(self.lib, self.lib.lagb_force_mult(self.h, self._p(e), self._p(v)))
        return v

    def force_mult_transpose(self, v):
        e = self.empty(self.P.ndofs_l2)
        _check(self.lib, self.lib.lagb_force_mult_transpose(self.h, self._p(v), self._p(e)))
        return e

    def qupdate(self, S, cfl=0.5, dt_est_in=float("inf")):
        out = C.c_double()
        _check(self.lib, self.lib.lagb_qupdate(self.h, self._p(S), cfl, dt_est_in, C.byref(out)))
        return out.value

    def pcg_vmass(self, comp, b, x=None, rel_tol=1e-8, max_iter=300):
        x = self.zeros(self.P.ndofs_h1) if x is None else x
        it = C.c_int32()
        _check(self.lib, self.lib.lagb_pcg_vmass(self.h, comp, self._p(b), self._p(x), rel_tol, max_iter, C.byref(it)))
        return x, it.value

    def pcg_vmass_all(self, rhs, x=None, rel_tol=1e-8, max_iter=300):
        x = self.zeros(self.P.h1_vsize) if x is None else x
        it = (C.c_int32 * 3)()
        _check(self.lib, self.lib.lagb_pcg_vmass_all(self.h, self._p(rhs), self._p(x), rel_tol, max_iter, it))
        return x, [int(it[c]) for c in range(self.P.dim)]

    def pcg_vmass_all_x0(self, rhs, rel_tol=1e-8, max_iter=300):
        """The batched solve from a zero initial guess (the reference's SolveVelocity): x is output only."""
        x = self.empty(self.P.h1_vsize)
        it = (C.c_int32 * 3)()
        _check(self.lib, self.lib.lagb_pcg_vmass_all_x0(self.h, self._p(rhs), self._p(x), rel_tol, max_iter, it))
        return x, [int(it[c]) for c in range(self.P.dim)]

    def cg_emass(self, b, rel_tol=1e-8, max_iter=300):
        x = self.empty(self.P.ndofs_l2)
        it = C.c_int32()
        _check(self.lib, self.lib.lagb_cg_emass(self.h, self._p(b), self._p(x), rel_tol, max_iter, C.byref(it)))
        return x, it.value

    def internal_energy(self, e):
        out = C.c_double()
        _check(self.lib, self.lib.lagb_internal_energy(self.h, self._p(e), C.byref(out)))
        return out.value

    def kinetic_energy(self, v):
        out = C.c_double()
        _check(self.lib, self.lib.lagb_kinetic_energy(self.h, self._p(v), C.byref(out)))
        return out.value

    def compute_density(self, x):
        """LagrangianHydroOperator::ComputeDensity (laghos_solver.cpp:542-563): L2 density on the mesh x."""
        rho = self.empty(self.P.ndofs_l2)
        _check(self.lib, self.lib.lagb_compute_density(self.h, self._p(x), self._p(rho)))
        return rho

    def taylor_source(self, x):
        e = self.empty(self.P.ndofs_l2)
        _check(self.lib, self.lib.lagb_taylor_source(self.h, self._p(x), self._p(e)))
        return e

    def qdata(self, which):
        """Copy of a QuadratureData array: 0 stressJinvT, 1 rho0DetJ0w, 2 Jac0inv, 3 mass D, 4 diagonal."""
        P = self.P
        n = {0: P.NE * P.NQ * P.dim * P.dim, 1: P.NE * P.NQ, 2: P.NE * P.NQ * P.dim * P.dim, 3: P.NE * P.NQ,
             4: P.ndofs_h1}[which]
        ptr = self.lib.lagb_qdata_ptr(self.h, which)
        out = self.empty(n)
        _check(self.lib, self.lib.lagb_vec_copy(self.h, self._p(out), C.c_void_p(ptr), n))
        self.sync()
        return out

    def set_sjit(self, t):
        ptr = self.lib.lagb_qdata_ptr(self.h, 0)
        _check(self.lib, self.lib.lagb_vec_copy(self.h, C.c_void_p(ptr), self._p(t), t.numel()))

    def timing(self):
        t = Timing()
        _check(self.lib, self.lib.lagb_timing_get(self.h, C.byref(t)))
        return t

    def timing_reset(self):
        _check(self.lib, self.lib.lagb_timing_reset(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.lagb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run(mesh="cube01_hex", rs=2, problem=1, ok=2, ot=1, oq=-1, blast_scale=None, impose_visc=False,
        ode_solver_type=4, t_final=0.6, max_tsteps=-1, cfl=0.5, cg_tol=1e-8, cg_max_iter=300,
        batched_pcg=True, kernel_variant=0, device=0, verbose=False, vis_steps=5, e2e_host_state=False,
        warmup_steps=0, rank=0, nranks=1, pgrid=(1, 1, 1), nccl_id=None, hist_cap=0, want_state=False,
        profile_mass=False, gfprint=False, visit=False, basename=None, v_error=False, check_exact_sedov=False, check=False, check_eps=0.0):
    """The reference driver's run (laghos.cpp main) through the C++ shim: lagb_laghos_run."""
    lib = load_library()
    dim = mesh_dim(mesh)
    if blast_scale is None:
        blast_scale = 1.0 / 2 ** dim
    o = RunOptions()
    lib.lagb_run_options_default(C.byref(o))
    mesh_b = mesh.encode()
    o.mesh = mesh_b
    o.rs, o.problem, o.ok, o.ot, o.oq = rs, problem, ok, ot, oq
    o.blast_scale, o.impose_visc = blast_scale, int(impose_visc)
    o.ode_solver_type, o.t_final, o.max_tsteps = ode_solver_type, t_final, max_tsteps
    o.cfl, o.cg_tol, o.cg_max_iter = cfl, cg_tol, cg_max_iter
    o.batched_pcg, o.kernel_variant, o.device = int(batched_pcg), kernel_variant, device
    o.verbose, o.vis_steps, o.e2e_host_state, o.warmup_steps = int(verbose), vis_steps, int(e2e_host_state), warmup_steps
    o.rank, o.nranks = rank, nranks
    o.profile_mass = int(profile_mass)
    o.gfprint, o.visit, o.v_error = int(gfprint), int(visit), int(v_error)
    o.check_exact_sedov = int(check_exact_sedov)
    o.check, o.check_eps = int(check), check_eps
    base_b = None if basename is None else str(basename).encode()   # kept alive until the call returns
    if base_b is not None:
        o.basename = base_b
    for d in range(3):
        o.pgrid[d] = pgrid[d]
    idbuf = None
    if nccl_id is not None:
        idbuf = C.create_string_buffer(bytes(nccl_id), 128)
        o.nccl_id = C.cast(idbuf, C.c_void_p)
    r = RunResult()
    hist = np.zeros(2 * max(hist_cap, 1), dtype=np.float64)
    S_out = None
    s_ptr = None
    if want_state:
        P = Problem(mesh, rs, problem, ok, ot, oq, blast_scale, impose_visc)
        S_out = np.zeros(P.s_size, dtype=np.float64)
        s_ptr = S_out.ctypes.data_as(c_double_p)
    rc = lib.lagb_laghos_run(C.byref(o), C.byref(r), hist.ctypes.data_as(c_double_p), hist_cap, s_ptr)
    if rc != 0:
        raise LagbError(f"lagb_laghos_run failed ({rc}): {lib.lagb_last_error().decode()}")
    out = dict(steps=r.steps, ti_last=r.ti_last, stages=r.stages, t=r.t, dt=r.dt, e_norm=r.e_norm,
               fom=list(r.fom), t_cgH1=r.timing.t_cgH1, t_cgL2=r.timing.t_cgL2, t_force=r.timing.t_force,
               t_qdata=r.timing.t_qdata, H1iter=r.timing.H1iter, L2iter=r.timing.L2iter,
               quad_tstep=r.timing.quad_tstep, wall_seconds=r.wall_seconds, device_seconds=r.device_seconds,
               mass_kernel_seconds=r.mass_kernel_seconds, mass_kernel_launches=r.mass_kernel_launches,
               mass_kernel_ncomp=r.mass_kernel_ncomp, work_mdof=r.work_mdof,
               energy_init=r.energy_init, energy_final=r.energy_final, v_err=list(r.v_err), density_l2_err=r.density_l2_err, checks=r.checks,
               h2d_bytes_per_step=r.h2d_bytes_per_step, d2h_bytes_per_step=r.d2h_bytes_per_step,
               kernel_launches=r.kernel_launches, ndofs_h1_global=r.ndofs_h1_global,
               ndofs_l2_global=r.ndofs_l2_global, ne_global=r.ne_global,
               hist=[(int(hist[2 * i]), float(hist[2 * i + 1])) for i in range(r.n_hist)])
    if want_state:
        out["S"] = S_out
    return out
