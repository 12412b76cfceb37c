"""INT8-slice (Ozaki) condensation on tcgen05 against the exact FP64 DMMA path and the oracle."""
import numpy as np
import pytest

from hiop_b200 import synth
from oracle import kkt_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from hiop_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _setup(ctx, P, mode):
    from hiop_b200.engine import KKTLinSysLowRank
    k = KKTLinSysLowRank(ctx, P.n, P.m_eq, P.m_ineq, max(P.l, 1))
    D = ctx.to_device
    J = D(P.J)
    T = {name: D(getattr(P, name)) for name in ("ixl", "ixu", "idl", "idu", "zl", "sxl", "zu", "sxu", "vl", "sdl", "vu", "sdu", "St", "Yt", "rx", "ryc", "ryd")}
    T["J"] = J
    k.set_patterns(T["ixl"], T["ixu"], T["idl"], T["idu"])
    k.set_jacobian(J[:P.m_eq], J[P.m_eq:])
    k.set_secant(P.sigma, T["St"] if P.l else None, T["Yt"] if P.l else None, P.L, P.D)
    k.set_condense_mode(mode)
    k.update(T["zl"], T["sxl"], T["zu"], T["sxu"], T["vl"], T["sdl"], T["vu"], T["sdu"])
    return k, T


# truncation after 6 + 7(S-1) bits relative to each row's largest entry; measured errors are ~10x below these bounds
TOL_N = {6: 2e-10, 7: 2e-12, 8: 5e-14}


@pytest.mark.parametrize("n,m,l", [(20000, 130, 6), (4099, 37, 3), (70000, 200, 6), (1000, 5, 0)])
@pytest.mark.parametrize("S", [6, 7, 8])
def test_ozaki_condense_matches_fp64(ctx, n, m, l, S):
    P = synth.make_qn_problem(n, m, l, seed=11 + n)
    k0, T0 = _setup(ctx, P, 0)
    k0.condense()
    N0 = k0.N()
    k1, T1 = _setup(ctx, P, S)
    k1.condense()
    N1 = k1.N()
    scale = np.sqrt(np.outer(np.diag(N0), np.diag(N0)))      # error model: relative to sqrt(N_ii N_jj)
    err = np.abs(N1 - N0) / scale
    assert np.array_equal(N1, N1.T)
    assert err.max() <= TOL_N[S], (S, err.max())
    # full solve: directions within the north-star tolerance
    dx0, dyc0, dyd0 = [ctx.zeros(s) for s in (P.n, P.m_eq, P.m_ineq)]
    dx1, dyc1, dyd1 = [ctx.zeros(s) for s in (P.n, P.m_eq, P.m_ineq)]
    assert k0.solveCompressed(ctx.to_device(P.rx), T0["ryc"], T0["ryd"], dx0, dyc0, dyd0)
    assert k1.solveCompressed(ctx.to_device(P.rx), T1["ryc"], T1["ryd"], dx1, dyc1, dyd1)
    ctx.sync()
    a, b = dx1.cpu().numpy(), dx0.cpu().numpy()
    assert np.abs(a - b).max() <= 1e-8 * np.abs(b).max()
    a, b = dyc1.cpu().numpy(), dyc0.cpu().numpy()
    assert np.abs(a - b).max() <= 1e-8 * max(1.0, np.abs(b).max())
    k0.close()
    k1.close()


def test_ozaki_rows_with_huge_dynamic_range(ctx):
    """rows whose entries span 12 decades and a row of exact zeros"""
    P = synth.make_qn_problem(30000, 40, 4, seed=5)
    r = np.random.default_rng(1)
    P.Jd[3] *= 10.0 ** r.uniform(-6, 6, P.n)
    P.Jd[2] = 0.0            # (the Dd^{-1} term keeps N SPD)
    P.Jc[5] *= 1e100
    P.Jd[7] *= 1e-100
    k0, _ = _setup(ctx, P, 0)
    k0.condense()
    N0 = k0.N()
    k1, _ = _setup(ctx, P, 8)
    k1.condense()
    N1 = k1.N()
    d = np.diag(N0).copy()
    d[d == 0] = 1.0
    err = np.abs(N1 - N0) / np.sqrt(np.outer(d, d))
    assert np.all(np.isfinite(N1))
    assert err.max() <= 1e-12, err.max()
    k0.close()
    k1.close()


def _oracle_N(P):
    Dx, DhInv, Dd, Dd_inv = ko.kkt_update(P.zl, P.sxl, P.zu, P.sxu, P.ixl, P.ixu, P.vl, P.sdl, P.vu, P.sdu, P.idl, P.idu, P.sigma)
    st = ko.QnState(P.Jc, P.Jd, DhInv, Dd_inv, P.St, P.Yt, P.L, P.D, P.sigma)
    N, W0, S1, Y1 = ko.condense(st)
    return N, DhInv


def test_oz8_against_oracle_at_scale_adversarial_scaling(ctx):
    """k_oz_gemm<8> against the ORACLE's restatement of symMatTimesInverseTimesMatTrans (src/Optimization/hiopHessianLowRank.cpp:549-630,
    the scalar triple loops :1079-1154; oracle/kkt_oracle.c is pinned to the compiled reference by tests/test_oracle_vs_ref.py) at a size
    with several K chunks (n = 200001 > 3 x 65535), an odd n, a 5 x 3 tile grid, DhInv spanning 16 decades and Jacobian rows spanning 12
    decades. Error model of the slicing: relative to the row scales, i.e. to sqrt(N_ii N_jj)."""
    n, m, l = 200001, 300, 4
    P = synth.make_qn_problem(n, m, l, seed=4242)
    r = np.random.default_rng(7)
    # Dx = zl/sxl in [1, 1e16]  ->  DhInv = 1/(sigma + Dx) spans 16 decades
    P.sxl[:] = 10.0 ** r.uniform(-8.0, 0.0, n)
    P.zl[:] = 10.0 ** r.uniform(0.0, 8.0, n)
    rowscale = 10.0 ** (12.0 * np.arange(m) / (m - 1) - 6.0)
    P.Jc *= rowscale[:P.m_eq, None]
    P.Jd *= rowscale[P.m_eq:, None]
    No, DhInv = _oracle_N(P)
    assert DhInv.max() / DhInv.min() > 1e15
    scale = np.sqrt(np.outer(np.abs(np.diag(No)), np.abs(np.diag(No))))
    for mode, tol in ((8, 2e-12), (0, 1e-12)):
        k, T = _setup(ctx, P, mode)
        k.condense()
        assert k.condense_mode_used() == mode
        Ng = k.N()
        err = (np.abs(Ng - No) / scale).max()
        assert err <= tol, (mode, err)
        k.close()


def test_auto_mode_uses_global_n_and_falls_back_to_fp64(ctx):
    """AUTO picks the kernel from the global column count (identical on every rank / world size) and, when the Cholesky of an
    int8-slice condensation breaks down, redoes it once with the exact FP64 kernel before reporting failure."""
    # (1) near-singular N: two identical equality rows + a tiny inequality regularisation. Exact FP64 keeps N (barely) positive definite or
    #     not -- either way AUTO must end where the FP64 mode ends, and count the retry when the int8 attempt failed.
    n, m, l = 40000, 70, 2
    P = synth.make_qn_problem(n, m, l, seed=99)
    P.Jc[1, :] = P.Jc[0, :] * (1.0 + 1e-15)
    outcomes = {}
    for mode in (0, -1):
        k, T = _setup(ctx, P, mode)
        try:
            k.condense()
            outcomes[mode] = ("ok", k.condense_mode_used(), k.fallback_count())
        except Exception as e:  # noqa: BLE001
            outcomes[mode] = ("fail", k.condense_mode_used(), k.fallback_count())
        k.close()
    assert outcomes[-1][0] == outcomes[0][0], outcomes
    if outcomes[-1][2] > 0:
        assert outcomes[-1][1] == 0           # the retry ran the FP64 kernel
    # (2) s, z down to 1e-10 (late interior-point iterates): AUTO (int8 slices) still delivers the direction within 1e-8
    P2 = synth.make_qn_problem(50000, 80, 4, seed=5)
    r = np.random.default_rng(3)
    P2.sxl[:] = 10.0 ** r.uniform(-10.0, 0.0, P2.n)
    P2.zl[:] = 10.0 ** r.uniform(-10.0, 0.0, P2.n)
    k, T = _setup(ctx, P2, -1)
    k.condense()
    assert k.condense_mode_used() == 8
    dx, dyc, dyd = [ctx.zeros(s) for s in (P2.n, P2.m_eq, P2.m_ineq)]
    assert k.solveCompressed(ctx.to_device(P2.rx), T["ryc"], T["ryd"], dx, dyc, dyd)
    k.check()
    ctx.sync()
    Dx, DhInv, Dd, Dd_inv = ko.kkt_update(P2.zl, P2.sxl, P2.zu, P2.sxu, P2.ixl, P2.ixu, P2.vl, P2.sdl, P2.vu, P2.sdu, P2.idl, P2.idu, P2.sigma)
    st = ko.QnState(P2.Jc, P2.Jd, DhInv, Dd_inv, P2.St, P2.Yt, P2.L, P2.D, P2.sigma)
    dxo, dyco, dydo, _ = ko.solve_compressed(st, P2.rx, P2.ryc, P2.ryd)
    assert np.abs(dx.cpu().numpy() - dxo).max() <= 1e-8 * np.abs(dxo).max()
    assert np.abs(dyd.cpu().numpy() - dydo).max() <= 1e-8 * max(1.0, np.abs(dydo).max())
    k.close()


@pytest.mark.parametrize("n,m,l", [(40000, 90, 4), (33001, 70, 3), (40960, 64, 0), (300, 66, 2)])
def test_fused_rowmax_sweep_delivers_the_same_direction(ctx, n, m, l):
    """solveCompressed with a PENDING int8-slice condensation takes J (H+Dx)^-1 rx from the row-maximum sweep (tdot - Z p) instead of a
    second pass over J; with the condensation already done it takes the two-pass route (hiopKKTLinSys.cpp:1146-1157). Same direction,
    and both agree with the oracle's solveCompressed."""
    P = synth.make_qn_problem(n, m, l, seed=n % 97)
    res = {}
    for fused in (True, False):
        k, T = _setup(ctx, P, 8)
        if not fused:
            k.condense()                              # condensation no longer pending -> unfused steps 1-2
        dx, dyc, dyd = [ctx.zeros(s) for s in (P.n, P.m_eq, P.m_ineq)]
        assert k.solveCompressed(ctx.to_device(P.rx), T["ryc"], T["ryd"], dx, dyc, dyd)
        k.check()
        ctx.sync()                                    # check() only waits when a condensation was still unchecked
        assert k.condense_mode_used() == 8
        res[fused] = [v.cpu().numpy().copy() for v in (dx, dyc, dyd)]
        # a second solve with another rhs on the same (valid) condensation must not reuse the dots of the first
        rx2 = ctx.to_device(P.rx[::-1].copy())
        dx2, dyc2, dyd2 = [ctx.zeros(s) for s in (P.n, P.m_eq, P.m_ineq)]
        assert k.solveCompressed(rx2, T["ryc"], T["ryd"], dx2, dyc2, dyd2)
        ctx.sync()
        res[(fused, 2)] = dx2.cpu().numpy().copy()
        k.close()
    for a, b in zip(res[True], res[False]):
        assert np.abs(a - b).max() <= 1e-10 * max(1.0, np.abs(b).max())
    assert np.abs(res[(True, 2)] - res[(False, 2)]).max() <= 1e-10 * max(1.0, np.abs(res[(False, 2)]).max())
    Dx, DhInv, Dd, Dd_inv = ko.kkt_update(P.zl, P.sxl, P.zu, P.sxu, P.ixl, P.ixu, P.vl, P.sdl, P.vu, P.sdu, P.idl, P.idu, P.sigma)
    st = ko.QnState(P.Jc, P.Jd, DhInv, Dd_inv, P.St, P.Yt, P.L, P.D, P.sigma)
    dxo, dyco, dydo, _ = ko.solve_compressed(st, P.rx, P.ryc, P.ryd)
    assert np.abs(res[True][0] - dxo).max() <= 1e-9 * max(1.0, np.abs(dxo).max())
    assert np.abs(res[True][2] - dydo).max() <= 1e-9 * max(1.0, np.abs(dydo).max())
