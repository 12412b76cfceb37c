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
