"""Debug aid: fused vs unfused solveCompressed vs the oracle on one small problem."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiop_b200 import synth
from hiop_b200.engine import Context, KKTLinSysLowRank
from oracle import kkt_oracle as ko

ctx = Context(0)
n, m, l = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (40960, 64, 0)
P = synth.make_qn_problem(n, m, l, seed=7)
Dx, DhInv, Dd, Dd_inv = ko.kkt_update(P.zl, P.sxl, P.zu, P.sxu, P.ixl, P.ixu, P.vl, P.sdl, P.vu, P.sdu, P.idl, P.idu, P.sigma)
st = ko.QnState(P.Jc, P.Jd, DhInv, Dd_inv, P.St, P.Yt, P.L, P.D, P.sigma)
dxo, dyco, dydo, _ = ko.solve_compressed(st, P.rx, P.ryc, P.ryd)
for mode in (8, 0):
    for fused in (True, False):
        k = KKTLinSysLowRank(ctx, P.n, P.m_eq, P.m_ineq, max(P.l, 1))
        D = ctx.to_device
        J = D(P.J)
        T = {name: D(getattr(P, name)) for name in ("ixl", "ixu", "idl", "idu", "zl", "sxl", "zu", "sxu", "vl", "sdl", "vu", "sdu", "St", "Yt", "rx", "ryc", "ryd")}
        k.set_patterns(T["ixl"], T["ixu"], T["idl"], T["idu"])
        k.set_jacobian(J[:P.m_eq], J[P.m_eq:])
        k.set_secant(P.sigma, T["St"] if P.l else None, T["Yt"] if P.l else None, P.L, P.D)
        k.set_condense_mode(mode)
        k.update(T["zl"], T["sxl"], T["zu"], T["sxu"], T["vl"], T["sdl"], T["vu"], T["sdu"])
        if not fused:
            k.condense()
        rx = D(P.rx)
        dx, dyc, dyd = [ctx.zeros(s) for s in (P.n, P.m_eq, P.m_ineq)]
        k.solveCompressed(rx, T["ryc"], T["ryd"], dx, dyc, dyd)
        k.check()
        ctx.sync()
        e = np.abs(dx.cpu().numpy() - dxo).max() / np.abs(dxo).max()
        ey = np.abs(dyd.cpu().numpy() - dydo).max() / max(1.0, np.abs(dydo).max())
        print(f"mode {mode} pending={fused}: used {k.condense_mode_used()} dx err {e:.3e} dyd err {ey:.3e}")
        k.close()
