"""hiopHessianLowRank::update on the device against the oracle restatement (oracle.kkt_oracle.SecantMemory,
pinned to the reference by tests/test_oracle_vs_ref.py::test_secant_update_matches_reference) over an iterate sequence that
covers the first call, appends, shifts and both skip rules, for every sigma rule; then the KKT solve that follows uses the
engine-owned memory."""
import numpy as np
import pytest

from hiop_b200 import synth
from oracle import kkt_oracle as ko
from test_gpu_parity import ctx  # noqa: F401

pytestmark = pytest.mark.gpu


def _mk(ctx, n, me, mi, lmax):
    from hiop_b200.engine import KKTLinSysLowRank
    k = KKTLinSysLowRank(ctx, n, me, mi, lmax)
    pat = [ctx.to_device(np.ones(n)), ctx.to_device(np.zeros(n)), ctx.to_device(np.ones(mi)), ctx.to_device(np.zeros(mi))]
    k.set_patterns(*pat)
    return k, pat


@pytest.mark.parametrize("strategy", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("n", [300, 1001])
def test_secant_sequence_against_oracle(ctx, strategy, n):
    me, mi, lmax = 4, 3, 3
    seq = synth.make_secant_sequence(n, me, mi, steps=8)
    k, pat = _mk(ctx, n, me, mi, lmax)
    k.secant_reset(1.0, strategy)
    mem = ko.SecantMemory(n, lmax, 1.0, strategy)
    for it in seq:
        J = ctx.to_device(np.vstack([it["Jc"], it["Jd"]]))
        k.set_jacobian(J[:me], J[me:])
        st = k.secant_update(ctx.to_device(it["x"]), ctx.to_device(it["grad_f"]), ctx.to_device(it["yc"]), ctx.to_device(it["yd"]))
        assert st == mem.update(it["x"], it["grad_f"], it["yc"], it["yd"], it["Jc"], it["Jd"])
        l, sigma, St, Yt, L, D = k.secant_state()
        assert l == mem.St.shape[0]
        np.testing.assert_array_equal(St, mem.St)                      # one subtraction per entry: same bits
        assert np.abs(Yt - mem.Yt).max(initial=0.0) <= 1e-13 * max(1.0, np.abs(mem.Yt).max(initial=0.0))
        assert np.abs(np.tril(L, -1) - np.tril(mem.L, -1)).max(initial=0.0) <= 1e-12
        assert np.abs(D - mem.D).max(initial=0.0) <= 1e-12
        assert abs(sigma - mem.sigma) <= 1e-12 * mem.sigma
    k.close()


def test_constant_jacobian_and_empty_memory(ctx):
    n, me, mi = 500, 2, 0
    seq = synth.make_secant_sequence(n, me, mi, steps=4)
    k, pat = _mk(ctx, n, me, mi, 0)                 # secant_memory_len = 0: only sigma is updated (hiopHessianLowRank.cpp:307)
    k.secant_reset(1.0, 1)
    mem = ko.SecantMemory(n, 0, 1.0, 1)
    J = ctx.to_device(np.vstack([seq[0]["Jc"], seq[0]["Jd"]]))
    k.set_jacobian(J[:me], J[me:])
    for it in seq:
        st = k.secant_update(ctx.to_device(it["x"]), ctx.to_device(it["grad_f"]), ctx.to_device(it["yc"]), ctx.to_device(it["yd"]), True)
        assert st == mem.update(it["x"], it["grad_f"], it["yc"], it["yd"], seq[0]["Jc"], seq[0]["Jd"])
        l, sigma, *_ = k.secant_state()
        assert l == 0 and abs(sigma - mem.sigma) <= 1e-12 * mem.sigma
    k.close()


def test_kkt_solve_uses_engine_owned_memory(ctx):
    """After device-side updates the condensed solve must equal the oracle's solve with the oracle's memory."""
    n, me, mi, lmax = 2000, 6, 5, 4
    seq = synth.make_secant_sequence(n, me, mi, steps=7, seed=5)
    P = synth.make_qn_problem(n, me + mi, 0, seed=11)              # iterate blocks / patterns / rhs of matching shape
    assert P.m_eq + P.m_ineq == me + mi
    me, mi = P.m_eq, P.m_ineq
    seq = synth.make_secant_sequence(n, me, mi, steps=7, seed=5)
    from hiop_b200.engine import KKTLinSysLowRank
    k = KKTLinSysLowRank(ctx, n, me, mi, lmax)
    T = {kk: ctx.to_device(getattr(P, kk)) for kk in ("ixl", "ixu", "idl", "idu", "sxl", "sxu", "zl", "zu", "sdl", "sdu", "vl", "vu")}
    k.set_patterns(T["ixl"], T["ixu"], T["idl"], T["idu"])
    k.secant_reset(1.0, 1)
    mem = ko.SecantMemory(n, lmax, 1.0, 1)
    for it in seq:
        J = ctx.to_device(np.vstack([it["Jc"], it["Jd"]]))
        k.set_jacobian(J[:me], J[me:])
        k.secant_update(ctx.to_device(it["x"]), ctx.to_device(it["grad_f"]), ctx.to_device(it["yc"]), ctx.to_device(it["yd"]))
        mem.update(it["x"], it["grad_f"], it["yc"], it["yd"], it["Jc"], it["Jd"])
    assert k.update(T["zl"], T["sxl"], T["zu"], T["sxu"], T["vl"], T["sdl"], T["vu"], T["sdu"])
    rx, ryc, ryd = ctx.to_device(P.rx), ctx.to_device(P.ryc), ctx.to_device(P.ryd)
    dx, dyc, dyd = ctx.zeros(n), ctx.zeros(me), ctx.zeros(mi)
    assert k.solveCompressed(rx, ryc, ryd, dx, dyc, dyd)
    ctx.sync()
    Dx, DhInv, Dd, Dd_inv = ko.kkt_update(P.zl, P.sxl, P.zu, P.sxu, P.ixl, P.ixu, P.vl, P.sdl, P.vu, P.sdu, P.idl, P.idu, mem.sigma)
    st = ko.QnState(seq[-1]["Jc"], seq[-1]["Jd"], DhInv, Dd_inv, mem.St, mem.Yt, mem.L, mem.D, mem.sigma)
    dxo, dyco, dydo, _ = ko.solve_compressed(st, P.rx, P.ryc, P.ryd)
    for a, b in ((dx.cpu().numpy(), dxo), (dyc.cpu().numpy(), dyco), (dyd.cpu().numpy(), dydo)):
        assert np.abs(a - b).max(initial=0.0) <= 1e-8 * max(1.0, np.abs(b).max(initial=0.0))
    k.close()
