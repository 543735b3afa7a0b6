#!/usr/bin/env python
"""Generate tests/golden/sedov_exact.json with the REFERENCE's own Sedov solution (sedov/sedov_sol.cpp compiled by
`make ref` into oracle/_ref/libsedov_ref.so -- only where /root/reference exists).  The vectors pin the product's
restatement (laghos_b200/csrc/host/sedov_exact.hpp) where the reference tree is absent (GPU box):
    python tools/make_sedov_golden.py"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    subprocess.check_call(["make", "ref"], cwd=ROOT)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsedov_ref.so"))
    dp = C.POINTER(C.c_double)
    lib.sedov_ref_eval.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                   dp, dp, dp, dp, dp]
    out = {"generated_by": "tools/make_sedov_golden.py from /root/reference/sedov/sedov_sol.cpp (SedovSol, omega = 0)",
           "cases": []}
    for dim in (2, 3):
        for gamma in (1.4, 5.0 / 3.0):
            for t, E0 in ((0.6, 1.0), (0.1, 0.25), (0.8, 2.0)):
                info = np.zeros(6)
                z = np.zeros(1)
                p = lambda a: a.ctypes.data_as(dp)
                assert lib.sedov_ref_eval(dim, gamma, 1.0, E0, 0.0, t, 0, p(z), p(z), p(z), p(z), p(info)) == 0
                r = np.concatenate([np.linspace(0.1, 0.999, 30) * info[1], np.array([1.0001, 1.2, 2.0]) * info[1]])
                rho, v, P = np.zeros_like(r), np.zeros_like(r), np.zeros_like(r)
                assert lib.sedov_ref_eval(dim, gamma, 1.0, E0, 0.0, t, r.size, p(r), p(rho), p(v), p(P), p(info)) == 0
                out["cases"].append(dict(dim=dim, gamma=gamma, t=t, blast_energy=E0, rho0=1.0,
                                         info=dict(zip(["alpha", "r2", "U", "rho2", "v2", "p2"], info.tolist())),
                                         r=r.tolist(), rho=rho.tolist(), v=v.tolist(), p=P.tolist()))
    path = os.path.join(ROOT, "tests", "golden", "sedov_exact.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
