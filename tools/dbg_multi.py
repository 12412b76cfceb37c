"""Debug aid (2 GPUs): which stage of the sharded solve differs between the ranks? python tools/dbg_multi.py"""
import os, sys, ctypes
import numpy as np
import torch
import torch.multiprocessing as mp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiop_b200 import synth, sharding


def worker(rank, world, uid, n, m, l, out):
    from hiop_b200.engine import Context, KKTLinSysLowRank, LinSolverSymDense
    torch.cuda.set_device(rank)
    ctx = Context(rank)
    ctx.init_comm(world, rank, uid)
    P = synth.make_qn_problem(n, m, l, seed=77)
    b, e = sharding.column_range(n, world, rank)
    sl = slice(b, e)
    D = ctx.to_device
    J = D(np.ascontiguousarray(P.J[:, sl]))
    # (1) J x with the all-reduce
    x = D(np.ascontiguousarray(P.rx[sl]))
    y = ctx.zeros(m)
    ctx.mat_times_vec(J, 0.0, y, 1.0, x)
    ctx.sync()
    out[f"jx{rank}"] = y.cpu().numpy()
    # (2) the same dense SPD solve on both devices
    A = np.random.default_rng(3).standard_normal((m, m))
    S = A @ A.T + m * np.eye(m)
    for mode, name in ((LinSolverSymDense.CHOLESKY, "chol"),):
        s = LinSolverSymDense(ctx, m, mode)
        s.set_matrix(D(np.triu(S)))
        s.matrixChanged()
        xb = D(P.ryc[:1].repeat(m) + np.arange(m))
        s.solve(xb)
        ctx.sync()
        out[f"{name}{rank}"] = xb.cpu().numpy()
        s.close()
    # (3) the sharded KKT solve, twice
    k = KKTLinSysLowRank(ctx, e - b, P.m_eq, P.m_ineq, max(l, 1))
    T = {name: D(np.ascontiguousarray(getattr(P, name)[sl])) for name in ("ixl", "ixu", "zl", "sxl", "zu", "sxu", "rx")}
    T.update({name: D(getattr(P, name)) for name in ("idl", "idu", "vl", "sdl", "vu", "sdu", "ryc", "ryd")})
    St, Yt = D(np.ascontiguousarray(P.St[:, sl])), D(np.ascontiguousarray(P.Yt[:, sl]))
    k.set_patterns(T["ixl"], T["ixu"], T["idl"], T["idu"])
    k.set_jacobian(J[:P.m_eq], J[P.m_eq:])
    k.set_secant(P.sigma, St if l else None, Yt if l else None, P.L, P.D)
    for rep in range(2):
        k.update(T["zl"], T["sxl"], T["zu"], T["sxu"], T["vl"], T["sdl"], T["vu"], T["sdu"])
        rx = D(np.ascontiguousarray(P.rx[sl]))
        dx, dyc, dyd = ctx.zeros(e - b), ctx.zeros(P.m_eq), ctx.zeros(P.m_ineq)
        k.solveCompressed(rx, T["ryc"], T["ryd"], dx, dyc, dyd)
        ctx.sync()
        out[f"dy{rep}{rank}"] = np.concatenate([dyc.cpu().numpy(), dyd.cpu().numpy()])
        out[f"N{rep}{rank}"] = k.N()
        out[f"stats{rep}{rank}"] = k.last_solve_stats()
    k.close()
    ctx.close()


if __name__ == "__main__":
    from hiop_b200 import _lib
    for (n, m, l) in ((9000, 140, 0), (9000, 100, 0), (9000, 140, 2)):
        buf = ctypes.create_string_buffer(128)
        _lib.check(_lib.lib().hb_comm_unique_id(buf), "uid")
        out = mp.Manager().dict()
        mp.spawn(worker, args=(2, buf.raw, n, m, l, out), nprocs=2, join=True)
        print(n, m, l, "Jx equal:", np.array_equal(out["jx0"], out["jx1"]), "chol equal:", np.array_equal(out["chol0"], out["chol1"]),
              "N equal:", np.array_equal(out["N00"], out["N01"]), "dy equal (1st, 2nd):", np.array_equal(out["dy00"], out["dy01"]),
              np.array_equal(out["dy10"], out["dy11"]), "same rank twice:", np.array_equal(out["dy00"], out["dy10"]),
              "max diff", np.abs(out["dy00"] - out["dy01"]).max(), out["stats00"], out["stats01"])
