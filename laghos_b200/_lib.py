"""ctypes loader and prototypes for liblaghos_b200.so (include/laghos_b200.h)."""
import ctypes as C
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "liblaghos_b200.so")
_lib = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int32)


class ProblemInfo(C.Structure):
    _fields_ = [("dim", C.c_int32), ("NE", C.c_int32), ("D1D", C.c_int32), ("L1D", C.c_int32),
                ("Q1D", C.c_int32), ("ND", C.c_int32), ("NL", C.c_int32), ("NQ", C.c_int32),
                ("nelem", C.c_int32 * 3), ("n1", C.c_int32 * 3),
                ("ndofs_h1", C.c_int64), ("ndofs_l2", C.c_int64),
                ("use_visc", C.c_int32), ("use_vort", C.c_int32), ("source", C.c_int32),
                ("ness", C.c_int32 * 3)]


class CtxDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("NE", C.c_int32), ("D1D", C.c_int32), ("L1D", C.c_int32), ("Q1D", C.c_int32),
                ("ndofs_h1", C.c_int64),
                ("h_h1_map", C.c_void_p), ("h_ess", C.c_void_p * 3), ("ness", C.c_int32 * 3),
                ("h_B", C.c_void_p), ("h_G", C.c_void_p), ("h_BL", C.c_void_p),
                ("h_qweights", C.c_void_p), ("h_gamma", C.c_void_p),
                ("use_visc", C.c_int32), ("use_vort", C.c_int32), ("device", C.c_int32),
                ("kernel_variant", C.c_int32), ("elem_grid", C.c_int32 * 3)]


class Timing(C.Structure):
    _fields_ = [("t_cgH1", C.c_double), ("t_cgL2", C.c_double), ("t_force", C.c_double), ("t_qdata", C.c_double),
                ("H1iter", C.c_int64), ("L2iter", C.c_int64), ("quad_tstep", C.c_int64)]


class RunOptions(C.Structure):
    _fields_ = [("mesh", C.c_char_p), ("rs", C.c_int), ("problem", C.c_int), ("ok", C.c_int), ("ot", C.c_int),
                ("oq", C.c_int), ("blast_scale", C.c_double), ("impose_visc", C.c_int),
                ("ode_solver_type", C.c_int), ("t_final", C.c_double), ("max_tsteps", C.c_int),
                ("cfl", C.c_double), ("cg_tol", C.c_double), ("cg_max_iter", C.c_int),
                ("batched_pcg", C.c_int), ("kernel_variant", C.c_int), ("device", C.c_int),
                ("verbose", C.c_int), ("vis_steps", C.c_int), ("e2e_host_state", C.c_int),
                ("warmup_steps", C.c_int), ("rank", C.c_int), ("nranks", C.c_int), ("pgrid", C.c_int * 3),
                ("nccl_id", C.c_void_p), ("profile_mass", C.c_int),
                ("gfprint", C.c_int), ("visit", C.c_int), ("basename", C.c_char_p),
                ("check_exact_sedov", C.c_int), ("v_error", C.c_int),
                ("check", C.c_int), ("check_eps", C.c_double)]


class RunResult(C.Structure):
    _fields_ = [("steps", C.c_int), ("ti_last", C.c_int), ("stages", C.c_int),
                ("t", C.c_double), ("dt", C.c_double), ("e_norm", C.c_double),
                ("fom", C.c_double * 5), ("timing", Timing),
                ("wall_seconds", C.c_double), ("device_seconds", C.c_double),
                ("h2d_bytes_per_step", C.c_int64), ("d2h_bytes_per_step", C.c_int64),
                ("kernel_launches", C.c_int64), ("n_hist", C.c_int),
                ("ndofs_h1_global", C.c_int64), ("ndofs_l2_global", C.c_int64), ("ne_global", C.c_int64),
                ("mass_kernel_seconds", C.c_double), ("mass_kernel_launches", C.c_int64),
                ("mass_kernel_ncomp", C.c_int64), ("work_mdof", C.c_double),
                ("energy_init", C.c_double), ("energy_final", C.c_double),
                ("v_err", C.c_double * 3), ("density_l2_err", C.c_double), ("checks", C.c_int)]


# every symbol include/laghos_b200.h declares (tests/test_abi_symbols.py checks the
# header against this list and against the built library)
SYMBOLS = """lagb_last_error lagb_kernel_launch_count lagb_problem_create lagb_problem_create_rect
lagb_problem_create_file lagb_problem_mesh_breaks
lagb_problem_create_part lagb_problem_nnbr lagb_problem_nbr lagb_problem_owner_mask
lagb_problem_destroy lagb_problem_get_info lagb_problem_h1_map lagb_problem_ess lagb_problem_S0
lagb_problem_rho0_gf lagb_problem_rho0_q lagb_problem_gamma lagb_problem_qweights lagb_problem_table
lagb_ctx_create lagb_ctx_destroy lagb_ctx_sync lagb_setup_qdata0 lagb_vmass_mult lagb_vmass_diag
lagb_emass_mult lagb_force_mult lagb_force_mult_transpose lagb_qupdate lagb_dt_est_set
lagb_qupdate_async lagb_dt_est_read lagb_pcg_vmass lagb_pcg_vmass_all lagb_cg_emass lagb_taylor_source
lagb_qdata_ptr lagb_qdata_h0 lagb_qdata_set_h0 lagb_dev_malloc lagb_dev_free lagb_memcpy_h2d
lagb_memcpy_h2d_async lagb_memcpy_d2h lagb_memcpy_h2d_bg lagb_memcpy_d2h_bg lagb_wait_copies lagb_host_alloc_pinned lagb_host_free_pinned lagb_vec_fill
lagb_vec_copy lagb_vec_axpby lagb_vec_dot lagb_nccl_unique_id lagb_ctx_comm_init lagb_allreduce_host
lagb_timing_get lagb_timing_reset lagb_stopwatch_start lagb_stopwatch_stop
lagb_profile_mass lagb_profile_mass_get lagb_vmass_mult_all lagb_tune_set lagb_internal_energy lagb_kinetic_energy
lagb_host_batch_plan_check lagb_compute_density lagb_pcg_vmass_all_x0
lagb_checks_entry lagb_checks_step lagb_problem_velocity_error lagb_sedov_exact_eval lagb_problem_sedov_density_error lagb_problem_write_mesh lagb_problem_write_field lagb_problem_write_print lagb_problem_write_visit""".split()


def load_library():
    """Load the CUDA library; fail loudly if it has not been built (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
            "laghos_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.lagb_last_error.restype = C.c_char_p
    lib.lagb_kernel_launch_count.restype = i64
    lib.lagb_problem_create.argtypes = [C.POINTER(vp), C.c_char_p, i32, i32, i32, i32, i32, dbl, i32]
    lib.lagb_problem_create_rect.argtypes = [C.POINTER(vp), i32, c_double_p, i32, c_double_p, i32, c_double_p, i32,
                                             i32, i32, i32, i32, i32, dbl, i32]
    lib.lagb_problem_create_file.argtypes = [C.POINTER(vp), C.c_char_p, i32, i32, i32, i32, i32, dbl, i32]
    lib.lagb_problem_mesh_breaks.argtypes = [vp, i32, C.POINTER(c_double_p), c_int_p]
    lib.lagb_problem_create_part.argtypes = [C.POINTER(vp), C.c_char_p, i32, i32, i32, i32, i32, dbl, i32, i32,
                                             C.POINTER(i32 * 3)]
    lib.lagb_problem_nnbr.argtypes = [vp]
    lib.lagb_problem_nbr.argtypes = [vp, i32, c_int_p, c_int_p, c_int_p, C.POINTER(c_int_p)]
    lib.lagb_problem_owner_mask.argtypes = [vp]
    lib.lagb_problem_owner_mask.restype = C.POINTER(C.c_uint8)
    lib.lagb_problem_destroy.argtypes = [vp]
    lib.lagb_problem_get_info.argtypes = [vp, C.POINTER(ProblemInfo)]
    for name, rt in [("lagb_problem_h1_map", c_int_p), ("lagb_problem_S0", c_double_p),
                     ("lagb_problem_rho0_gf", c_double_p), ("lagb_problem_rho0_q", c_double_p),
                     ("lagb_problem_gamma", c_double_p), ("lagb_problem_qweights", c_double_p)]:
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = rt
    lib.lagb_problem_ess.argtypes = [vp, i32]
    lib.lagb_problem_ess.restype = c_int_p
    lib.lagb_problem_table.argtypes = [vp, i32]
    lib.lagb_problem_table.restype = c_double_p
    lib.lagb_problem_velocity_error.argtypes = [vp, vp, c_double_p]
    lib.lagb_checks_entry.argtypes = [i32, i32, i32, c_int_p, c_double_p]
    lib.lagb_checks_step.argtypes = [i32, i32, i32, dbl, dbl, c_int_p]
    lib.lagb_sedov_exact_eval.argtypes = [i32, dbl, dbl, dbl, dbl, dbl, i32, vp, vp, vp, vp, vp]
    lib.lagb_problem_sedov_density_error.argtypes = [vp, vp, vp, dbl, dbl, dbl, dbl, c_double_p]
    lib.lagb_problem_write_mesh.argtypes = [vp, vp, C.c_char_p, i32]
    lib.lagb_problem_write_field.argtypes = [vp, i32, i32, vp, C.c_char_p, i32]
    lib.lagb_problem_write_print.argtypes = [vp, C.c_char_p, i32, vp, vp, i32]
    lib.lagb_problem_write_visit.argtypes = [vp, C.c_char_p, i32, dbl, dbl, i32, i32, vp, vp, i32]
    lib.lagb_ctx_create.argtypes = [C.POINTER(vp), C.POINTER(CtxDesc), vp]
    lib.lagb_ctx_destroy.argtypes = [vp]
    lib.lagb_ctx_sync.argtypes = [vp]
    lib.lagb_setup_qdata0.argtypes = [vp, vp, vp, vp, i64, c_double_p]
    lib.lagb_vmass_mult.argtypes = [vp, i32, vp, vp]
    lib.lagb_vmass_diag.argtypes = [vp, vp]
    lib.lagb_vmass_mult_all.argtypes = [vp, vp, vp]
    lib.lagb_tune_set.argtypes = [vp, i32, i32]
    lib.lagb_host_batch_plan_check.argtypes = [c_int_p, i32, i32, i64, c_int_p, i32, C.POINTER(i64)]
    lib.lagb_emass_mult.argtypes = [vp, vp, vp]
    lib.lagb_force_mult.argtypes = [vp, vp, vp]
    lib.lagb_force_mult_transpose.argtypes = [vp, vp, vp]
    lib.lagb_qupdate.argtypes = [vp, vp, dbl, dbl, c_double_p]
    lib.lagb_dt_est_set.argtypes = [vp, dbl]
    lib.lagb_qupdate_async.argtypes = [vp, vp, dbl]
    lib.lagb_dt_est_read.argtypes = [vp, c_double_p]
    lib.lagb_pcg_vmass.argtypes = [vp, i32, vp, vp, dbl, i32, c_int_p]
    lib.lagb_pcg_vmass_all.argtypes = [vp, vp, vp, dbl, i32, c_int_p]
    lib.lagb_pcg_vmass_all_x0.argtypes = [vp, vp, vp, dbl, i32, c_int_p]
    lib.lagb_cg_emass.argtypes = [vp, vp, vp, dbl, i32, c_int_p]
    lib.lagb_taylor_source.argtypes = [vp, vp, vp]
    lib.lagb_compute_density.argtypes = [vp, vp, vp]
    lib.lagb_internal_energy.argtypes = [vp, vp, c_double_p]
    lib.lagb_kinetic_energy.argtypes = [vp, vp, c_double_p]
    lib.lagb_qdata_ptr.argtypes = [vp, i32]
    lib.lagb_qdata_ptr.restype = vp
    lib.lagb_qdata_h0.argtypes = [vp]
    lib.lagb_qdata_h0.restype = dbl
    lib.lagb_qdata_set_h0.argtypes = [vp, dbl]
    lib.lagb_vec_fill.argtypes = [vp, vp, dbl, i64]
    lib.lagb_vec_copy.argtypes = [vp, vp, vp, i64]
    lib.lagb_vec_axpby.argtypes = [vp, vp, dbl, vp, dbl, vp, i64]
    lib.lagb_vec_dot.argtypes = [vp, vp, vp, i64, c_double_p]
    lib.lagb_nccl_unique_id.argtypes = [C.c_char_p]
    lib.lagb_allreduce_host.argtypes = [vp, c_double_p, i32, i32]
    lib.lagb_timing_get.argtypes = [vp, C.POINTER(Timing)]
    lib.lagb_timing_reset.argtypes = [vp]
    lib.lagb_stopwatch_start.argtypes = [vp]
    lib.lagb_stopwatch_stop.argtypes = [vp, c_double_p]
    lib.lagb_profile_mass.argtypes = [vp, i32]
    lib.lagb_profile_mass_get.argtypes = [vp, c_double_p, C.POINTER(i64)]
    lib.lagb_laghos_run.argtypes = [C.POINTER(RunOptions), C.POINTER(RunResult), c_double_p, i32, c_double_p]
    lib.lagb_run_options_default.argtypes = [C.POINTER(RunOptions)]
    _lib = lib
    return lib
