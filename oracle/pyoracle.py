"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product (laghos_b200/) never does.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None
dp = C.POINTER(C.c_double)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `make oracle/_build/liboracle.so`")
        lib = C.CDLL(LIB_PATH)
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                   C.c_double, C.c_double, C.c_int, C.c_int]
        lib.orc_create_part.restype = C.c_void_p
        lib.orc_create_part.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                        C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int * 3)]
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_info.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        lib.orc_get_S0.argtypes = [C.c_void_p, dp]
        lib.orc_h0.argtypes = [C.c_void_p]
        lib.orc_h0.restype = C.c_double
        lib.orc_qdata.argtypes = [C.c_void_p, C.c_int, dp]
        lib.orc_qdata.restype = C.c_longlong
        lib.orc_set_sjit.argtypes = [C.c_void_p, dp]
        lib.orc_vmass_mult.argtypes = [C.c_void_p, C.c_int, dp, dp]
        lib.orc_emass_mult.argtypes = [C.c_void_p, dp, dp]
        lib.orc_force_mult.argtypes = [C.c_void_p, dp, dp]
        lib.orc_force_mult_t.argtypes = [C.c_void_p, dp, dp]
        lib.orc_qupdate.argtypes = [C.c_void_p, dp, C.c_double]
        lib.orc_qupdate.restype = C.c_double
        lib.orc_pcg_vmass.argtypes = [C.c_void_p, C.c_int, dp, dp]
        lib.orc_cg_emass.argtypes = [C.c_void_p, dp, dp]
        lib.orc_taylor_source.argtypes = [C.c_void_p, dp, dp]
        lib.orc_mult.argtypes = [C.c_void_p, dp, dp]
        lib.orc_internal_energy.argtypes = [C.c_void_p, dp]
        lib.orc_internal_energy.restype = C.c_double
        lib.orc_kinetic_energy.argtypes = [C.c_void_p, dp]
        lib.orc_kinetic_energy.restype = C.c_double
        lib.orc_compute_density.argtypes = [C.c_void_p, dp, dp]
        lib.orc_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                dp, dp, C.c_int, dp]
        _lib = lib
    return _lib


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(dp)


def _dim(mesh):
    return 2 if mesh in ("square01_quad", "rectangle01_quad", "square_gresho", "rt2D") else 3


class Oracle:
    def __init__(self, mesh="cube01_hex", rs=0, problem=1, ok=2, ot=1, oq=-1, blast_scale=None, impose_visc=False,
                 cfl=0.5, cg_tol=1e-8, cg_max_iter=300, nthreads=1, rank=0, pgrid=None):
        self.lib = load()
        if blast_scale is None:
            blast_scale = 1.0 / 2 ** _dim(mesh)
        if pgrid is None:
            self.h = self.lib.orc_create(mesh.encode(), rs, problem, ok, ot, oq, blast_scale, int(impose_visc),
                                         cfl, cg_tol, cg_max_iter, nthreads)
        else:
            pg = (C.c_int * 3)(*pgrid)
            self.h = self.lib.orc_create_part(mesh.encode(), rs, problem, ok, ot, oq, blast_scale, int(impose_visc),
                                              cfl, cg_tol, cg_max_iter, nthreads, rank, C.byref(pg))
        if not self.h:
            raise RuntimeError("orc_create failed")
        info = (C.c_longlong * 10)()
        self.lib.orc_info(self.h, info)
        (self.dim, self.NE, self.D1D, self.L1D, self.Q1D, self.ND, self.NL, self.NQ, self.ndofs_h1,
         self.ndofs_l2) = [int(v) for v in info]
        self.h1_vsize = self.dim * self.ndofs_h1
        self.s_size = 2 * self.h1_vsize + self.ndofs_l2

    def __del__(self):
        try:
            if self.h:
                self.lib.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def S0(self):
        S = np.zeros(self.s_size)
        self.lib.orc_get_S0(self.h, _p(S))
        return S

    @property
    def h0(self):
        return self.lib.orc_h0(self.h)

    def qdata(self, which):
        n = self.lib.orc_qdata(self.h, which, None)
        out = np.zeros(n)
        self.lib.orc_qdata(self.h, which, _p(out))
        return out

    def set_sjit(self, a):
        self.lib.orc_set_sjit(self.h, _p(np.ascontiguousarray(a)))

    def vmass_mult(self, x, comp=-1):
        y = np.zeros(self.ndofs_h1)
        self.lib.orc_vmass_mult(self.h, comp, _p(x), _p(y))
        return y

    def emass_mult(self, x):
        y = np.zeros(self.ndofs_l2)
        self.lib.orc_emass_mult(self.h, _p(x), _p(y))
        return y

    def force_mult(self, e):
        v = np.zeros(self.h1_vsize)
        self.lib.orc_force_mult(self.h, _p(e), _p(v))
        return v

    def force_mult_transpose(self, v):
        e = np.zeros(self.ndofs_l2)
        self.lib.orc_force_mult_t(self.h, _p(v), _p(e))
        return e

    def qupdate(self, S, dt_in=float("inf")):
        return self.lib.orc_qupdate(self.h, _p(S), dt_in)

    def pcg_vmass(self, comp, b, x=None):
        x = np.zeros(self.ndofs_h1) if x is None else x
        it = self.lib.orc_pcg_vmass(self.h, comp, _p(b), _p(x))
        return x, it

    def cg_emass(self, b):
        x = np.zeros(self.ndofs_l2)
        it = self.lib.orc_cg_emass(self.h, _p(b), _p(x))
        return x, it

    def taylor_source(self, x):
        e = np.zeros(self.ndofs_l2)
        self.lib.orc_taylor_source(self.h, _p(x), _p(e))
        return e

    def internal_energy(self, e):
        return self.lib.orc_internal_energy(self.h, _p(np.ascontiguousarray(e)))

    def kinetic_energy(self, v):
        return self.lib.orc_kinetic_energy(self.h, _p(np.ascontiguousarray(v)))

    def compute_density(self, x):
        rho = np.zeros(self.ndofs_l2)
        self.lib.orc_compute_density(self.h, _p(np.ascontiguousarray(x)), _p(rho))
        return rho

    def mult(self, S):
        d = np.zeros(self.s_size)
        self.lib.orc_mult(self.h, _p(S), _p(d))
        return d


def run(mesh="cube01_hex", rs=2, problem=1, ok=2, ot=1, oq=-1, blast_scale=None, impose_visc=False,
        ode_solver_type=4, t_final=0.6, max_tsteps=-1, cfl=0.5, cg_tol=1e-8, cg_max_iter=300, nthreads=1,
        hist_cap=100000, want_state=False):
    lib = load()
    if blast_scale is None:
        blast_scale = 1.0 / 2 ** _dim(mesh)
    out = np.zeros(32)
    hist = np.zeros(2 * max(1, hist_cap))
    S = None
    if want_state:
        o = Oracle(mesh, rs, problem, ok, ot, oq, blast_scale, impose_visc)
        S = np.zeros(o.s_size)
    rc = lib.orc_run(mesh.encode(), rs, problem, ok, ot, oq, blast_scale, int(impose_visc), ode_solver_type,
                     t_final, max_tsteps, cfl, cg_tol, cg_max_iter, nthreads, _p(out), _p(hist), hist_cap,
                     _p(S) if S is not None else None)
    if rc != 0:
        raise RuntimeError(f"orc_run failed: {rc}")
    n = int(out[18])
    res = dict(steps=int(out[0]), ti_last=int(out[1]), t=out[2], dt=out[3], e_norm=out[4], fom=list(out[5:10]),
               t_cgH1=out[10], t_cgL2=out[11], t_force=out[12], t_qdata=out[13], H1iter=int(out[14]),
               L2iter=int(out[15]), quad_tstep=int(out[16]), stages=int(out[17]),
               energy_init=out[19], energy_final=out[20],
               hist=[(int(hist[2 * i]), float(hist[2 * i + 1])) for i in range(n)])
    if want_state:
        res["S"] = S
    return res
