#!/usr/bin/env python
"""Launches and microseconds per BiCGStab iteration of hb_lowrank_compute_directions_w_ir (device-side outer refinement,
hiopKKTLinSys::compute_directions_w_IR src/Optimization/hiopKKTLinSys.cpp:909-960) in the launch-latency regime of the bundled drivers.
The tolerance is set below reach so that all `maxit` iterations run. HIOPB200_SO=<other build> gives the before/after pair.
Usage: python tools/bench_krylov.py [n m l maxit]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiop_b200 import synth  # noqa: E402
from hiop_b200.engine import Context, KKTLinSysLowRank  # noqa: E402
from oracle import kkt_oracle as ko  # noqa: E402  (names of the residual / direction blocks only)


def main():
    a = [int(x) for x in sys.argv[1:]]
    n, m, l, maxit = (a + [5000, 4, 6, 8])[:4] if len(a) < 4 else a[:4]
    P = synth.make_qn_problem(n, m, l, seed=9)
    ctx = Context(0)
    k = KKTLinSysLowRank(ctx, P.n, P.m_eq, P.m_ineq, max(l, 1))
    D = ctx.to_device
    J = D(P.J)
    T = {name: D(getattr(P, name)) for name in ("ixl", "ixu", "idl", "idu", "zl", "sxl", "zu", "sxu", "vl", "sdl", "vu", "sdu", "St", "Yt")}
    k.set_patterns(T["ixl"], T["ixu"], T["idl"], T["idu"])
    k.set_jacobian(J[:P.m_eq], J[P.m_eq:])
    k.set_secant(P.sigma, T["St"], T["Yt"], P.L, P.D)
    k.update(T["zl"], T["sxl"], T["zu"], T["sxu"], T["vl"], T["sdl"], T["vu"], T["sdu"])
    k.condense()
    res = {kk: D(P.res[kk]) for kk in ko.RES_NAMES}
    sizes = dict(x=P.n, d=P.m_ineq, yc=P.m_eq, yd=P.m_ineq, sxl=P.n, sxu=P.n, sdl=P.m_ineq, sdu=P.m_ineq, zl=P.n, zu=P.n, vl=P.m_ineq, vu=P.m_ineq)
    dirs = {kk: ctx.zeros(sizes[kk]) for kk in ko.DIR_NAMES}
    best, launches, its = 1e30, 0, 0
    for rep in range(5):
        ctx.sync()
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        ok, info = k.compute_directions_w_IR(res, dirs, mu=1e-30, maxit=maxit)   # tol = mu * 1e-2: unreachable -> maxit iterations
        ctx.sync()
        dt = time.perf_counter() - t0
        if dt < best:
            best, launches, its = dt, ctx.launch_count() - l0, info[1]
    print(f"n={n} m={m} l={l}: {its} BiCGStab iterations in {best * 1e3:.3f} ms = {best * 1e6 / max(its, 0.5):.1f} us/iteration, "
          f"{launches} launches = {launches / max(its, 0.5):.1f} per iteration, flag {info[0]} (library {os.environ.get('HIOPB200_SO', 'default')})")
    k.close()
    ctx.close()


if __name__ == "__main__":
    main()
