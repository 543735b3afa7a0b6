#!/usr/bin/env python
"""Per-operator timings on one GPU (CUDA events, L2-cold: every operator streams > L2 of data).

    python tools/microbench.py [--rs 5] [--ok 3] [--reps 10]

Prints one line per operator: average microseconds, algorithmic GB, GB/s and
fraction of the measured HBM peak.  Used to pick launch variants (lagb_tune_set).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rs", type=int, default=5)
    ap.add_argument("--ok", type=int, default=3)
    ap.add_argument("--mesh", default="cube01_hex")
    ap.add_argument("--problem", type=int, default=1)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--no-pcg", action="store_true")
    ap.add_argument("--mass-variants", default="0,1,2,3,4")
    ap.add_argument("--force-variants", default="0,1,2,3,4")
    ap.add_argument("--q-variants", default="0,1,2")
    ap.add_argument("--mass1-variants", default="0")
    ap.add_argument("--brick-variants", default="0,1,2,3,4")
    ap.add_argument("--brick-shapes", default="0,1,2")
    ap.add_argument("--only", default="", help="comma list of sections: q,force,mass,l2,pcg (default all)")
    args = ap.parse_args()
    import numpy as np
    import torch
    from laghos_b200.api import Problem, Context
    peak = 6488.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    P = Problem(args.mesh, args.rs, args.problem, args.ok, args.ok - 1)
    c = Context(P)
    NE, NQ, nd, nl = P.NE, P.NQ, P.ndofs_h1, P.ndofs_l2
    rng = np.random.default_rng(1)
    S = P.S0.copy()
    nv = P.h1_vsize
    S[nv:2 * nv] = 0.01 * rng.uniform(-1, 1, nv)
    S[2 * nv:] = rng.uniform(0.5, 1.5, nl)
    dS = c.dev(S)
    v = c.dev(rng.uniform(-1, 1, nv))
    e = c.dev(rng.uniform(0.5, 1.5, nl))
    x1 = c.dev(rng.uniform(-1, 1, nd))
    yv = c.empty(nv)
    c.qupdate(dS)

    class Unavailable(Exception):
        pass

    def timeit(fn, reps=args.reps):
        for _ in range(2):
            rc = fn()
            if isinstance(rc, int) and rc != 0:
                raise Unavailable(c.lib.lagb_last_error().decode())
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        return 1e3 * ts[len(ts) // 2]

    rows = []

    def report(name, fn_or_us, gbytes, out=None):
        if callable(fn_or_us):
            try:
                us = timeit(fn_or_us)
            except Unavailable as ex:
                print(f"{name:38s} unavailable: {ex}", flush=True)
                return
        else:
            us = fn_or_us
        bw = gbytes / (us * 1e-6)
        rows.append((name, us, gbytes, bw, bw / peak))
        chk = "" if out is None else f"  checksum {float(out.double().abs().sum()):.15e}"
        print(f"{name:38s} {us:10.1f} us  {gbytes:7.3f} GB  {bw:8.1f} GB/s  {100 * bw / peak:5.1f}% of {peak:.0f}{chk}", flush=True)

    dim = P.dim
    only = set(args.only.split(",")) if args.only else {"q", "force", "mass", "l2", "pcg"}
    for var in ([int(s) for s in args.q_variants.split(",")] if "q" in only else []):
        c.tune(2, var)
        report(f"qupdate (fused) variant {var}", (lambda: c.lib.lagb_qupdate_async(c.h, c._p(dS), 0.5)),
               8e-9 * (2 * dim * nd + nl + NE * NQ * (1 + 2 * dim * dim)), c.qdata(0))
    c.tune(2, 0)
    ye = c.empty(nl)
    for var in ([int(s) for s in args.force_variants.split(",")] if "force" in only else []):
        c.tune(1, var)
        report(f"force_mult variant {var}", (lambda: c.lib.lagb_force_mult(c.h, c._p(e), c._p(yv))),
               8e-9 * (dim * dim * NE * NQ + nl + dim * nd), yv)
        report(f"force_mult_transpose variant {var}", (lambda: c.lib.lagb_force_mult_transpose(c.h, c._p(v), c._p(ye))),
               8e-9 * (dim * dim * NE * NQ + nl + dim * nd), ye)
    c.tune(1, 0)
    y1 = c.empty(nd)
    # brick schedule (no atomics): key 6 = 3 second kernel / 2 first kernel, key 7 = brick shape, key 4 = launch variant,
    # key 5 = 1 disables programmatic dependent launch
    for path, shape in ([(4, int(s)) for s in args.brick_shapes.split(",")] + [(3, int(s)) for s in args.brick_shapes.split(",")] + [(2, 0)] if "mass" in only else []):
        c.tune(6, path)
        c.tune(7, shape)
        for var in [int(s) for s in args.brick_variants.split(",")]:
            if path == 2 and var > 2:
                continue
            c.tune(4, var)
            for pdl_off in ((0, 1) if path < 4 else (0,)):
                c.tune(5, pdl_off)
                report(f"vmass_mult_all brick{path - 1} shape {shape} variant {var} pdl {1 - pdl_off}",
                       timeit(lambda: c.lib.lagb_vmass_mult_all(c.h, c._p(v), c._p(yv))), 8e-9 * (NE * NQ + 2 * dim * nd), yv)
            c.tune(5, 0)
            report(f"vmass_mult (1 comp) brick{path - 1} shape {shape} variant {var}",
                   timeit(lambda: c.lib.lagb_vmass_mult(c.h, -1, c._p(x1), c._p(y1))), 8e-9 * (NE * NQ + 2 * nd), y1)
    c.tune(7, 0)
    c.tune(4, 0)
    c.tune(6, 1)   # legacy atomic-scatter kernels
    for var in ([int(s) for s in args.mass1_variants.split(",")] if "mass" in only else []):
        c.tune(3, var)
        report(f"vmass_mult (1 comp) variant {var}", (lambda: c.lib.lagb_vmass_mult(c.h, -1, c._p(x1), c._p(y1))),
               8e-9 * (NE * NQ + 2 * nd), y1)
    c.tune(3, 0)
    for var in ([int(s) for s in args.mass_variants.split(",")] if "mass" in only else []):
        c.tune(0, var)
        report(f"vmass_mult_all (3 comp) variant {var}", (lambda: c.lib.lagb_vmass_mult_all(c.h, c._p(v), c._p(yv))),
               8e-9 * (NE * NQ + 2 * dim * nd), yv)
    c.tune(0, 0)
    c.tune(6, 0)
    if "l2" in only:
        report("emass_mult (L2)", timeit(lambda: c.lib.lagb_emass_mult(c.h, c._p(e), c._p(ye))), 8e-9 * (NE * NQ + 2 * nl))
    if args.no_pcg or "pcg" not in only:
        c.close()
        return
    b = c.dev(rng.uniform(-1, 1, nv))
    xs = c.zeros(nv)

    def pcg():
        xs.zero_()
        c.pcg_vmass_all(b, xs)

    for path in (3, 2, 1):
        c.tune(6, path)
        us = timeit(pcg, 3)
        _, its = c.pcg_vmass_all(b, c.zeros(nv))
        nit = max(its)
        print(f"pcg_vmass_all ({['', 'legacy', 'brick1', 'brick2'][path]}): {us:.1f} us for {its} iterations -> {us / (nit + 1):.1f} us per iteration", flush=True)
    c.tune(6, 0)
    bl = c.dev(rng.uniform(-1, 1, nl))
    us = timeit(lambda: c.cg_emass(bl), 3)
    _, it2 = c.cg_emass(bl)
    print(f"cg_emass: {us:.1f} us for {it2} iterations -> {us / max(it2, 1):.1f} us per iteration", flush=True)
    c.close()


if __name__ == "__main__":
    main()
