"""Large-N path of the B1 solver (hb_dense_big.cu): look-ahead Cholesky / no-pivot LDL^T, 128 x 128 diagonal-block inverses, blocked
multi-CTA solves -- hiopLinSolverSymDense{Lapack,MagmaNopiv}::matrixChanged / solve (src/LinAlg/hiopLinSolverSymDenseLapack.hpp:75-192,
src/LinAlg/hiopLinSolverSymDenseMagma.cpp:324-476). Checked against LAPACK through numpy (the oracle's factor/solve are the same calls)."""
import numpy as np
import pytest

from hiop_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from hiop_b200.engine import Context
    c = Context(0)
    yield c
    c.close()


def _factor_solve(ctx, K, mode, nrhs=1, seed=0):
    from hiop_b200.engine import LinSolverSymDense
    N = K.shape[0]
    s = LinSolverSymDense(ctx, N, mode)
    s.set_matrix(ctx.to_device(np.triu(K) + np.tril(np.full((N, N), np.nan), -1)))  # the lower part must never be read
    ret = s.matrixChanged()
    rhs = np.random.default_rng(seed).standard_normal((nrhs, N))
    x = ctx.to_device(rhs.copy())
    ok = s.solve(x) if ret >= 0 else False
    ctx.sync()
    xs = x.cpu().numpy()
    s.close()
    return ret, ok, rhs, xs


@pytest.mark.parametrize("N", [130, 300, 1024, 1100, 2501, 4096])
def test_cholesky_large(ctx, N):
    from hiop_b200.engine import LinSolverSymDense
    r = np.random.default_rng(N)
    A = r.standard_normal((N, N // 2 + 1))
    S = A @ A.T + np.diag(r.uniform(0.5, 2.0, N))
    ret, ok, rhs, xs = _factor_solve(ctx, S, LinSolverSymDense.CHOLESKY, nrhs=2, seed=N)
    assert ret == 0 and ok
    ref = np.linalg.solve(S, rhs.T).T
    assert np.abs(xs - ref).max() <= 1e-9 * np.abs(ref).max()
    assert np.abs(S @ xs.T - rhs.T).max() <= 1e-9 * np.abs(rhs).max()


@pytest.mark.parametrize("nx,m", [(200, 57), (300, 213), (700, 324), (1500, 1001), (2200, 900)])
def test_nopiv_ldl_large(ctx, nx, m):
    from hiop_b200.engine import LinSolverSymDense
    K = synth.make_kkt_like(nx, m, seed=nx + m)
    ret, ok, rhs, xs = _factor_solve(ctx, K, LinSolverSymDense.NOPIV, nrhs=1, seed=m)
    assert ret == m and ok      # quasi-definite: same inertia as Bunch-Kaufman (magma nopiv mode)
    ref = np.linalg.solve(K, rhs.T).T
    assert np.abs(xs - ref).max() <= 1e-8 * np.abs(ref).max()


def test_large_breakdowns(ctx):
    from hiop_b200.engine import LinSolverSymDense
    N = 1300
    r = np.random.default_rng(5)
    A = r.standard_normal((N, N))
    S = A @ A.T + N * np.eye(N)
    S[700, 700] = -1.0                     # not SPD: leading minor 701 fails
    ret, ok, _, _ = _factor_solve(ctx, S, LinSolverSymDense.CHOLESKY)
    assert ret == -1
    K = synth.make_kkt_like(400, 100, seed=3)
    K[0, :] = 0.0
    K[:, 0] = 0.0                          # zero pivot in the first column
    ret, ok, _, _ = _factor_solve(ctx, K, LinSolverSymDense.NOPIV)
    assert ret == -1


def test_factor_is_reproducible(ctx):
    """two-stream look-ahead must not change the bits of the result (fixed summation orders, no atomics on data)"""
    from hiop_b200.engine import LinSolverSymDense
    K = synth.make_kkt_like(900, 400, seed=8)
    outs = [_factor_solve(ctx, K, LinSolverSymDense.NOPIV, seed=1)[3] for _ in range(3)]
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("N,nneg", [(400, 150), (513, 256), (1000, 333), (1537, 700)])
def test_bunch_kaufman_cluster_general_indefinite(ctx, N, nneg):
    """cluster panel kernel (hb_bk_cluster.cu): tiny diagonal -> 2x2 pivots and interchanges far outside the panel; the inertia must equal
    the eigenvalue count (DSYTRF + the dsidi rule of hiopLinSolverSymDenseLapack.hpp:127-167) and the solve must match LAPACK's"""
    from hiop_b200.engine import LinSolverSymDense
    M = synth.make_symmetric_indefinite(N, nneg, seed=N)
    M[np.diag_indices(N)] *= 1e-6
    ev = np.linalg.eigvalsh(M)
    ret, ok, rhs, xs = _factor_solve(ctx, M, LinSolverSymDense.BUNCH_KAUFMAN, nrhs=2, seed=N)
    assert ret == int((ev < 0).sum()) and ok
    ref = np.linalg.solve(M, rhs.T).T
    assert np.abs(M @ xs.T - rhs.T).max() <= 1e-9 * np.abs(rhs).max() * max(1.0, np.linalg.cond(M) * 1e-3)
    assert np.abs(xs - ref).max() <= 1e-7 * np.abs(ref).max()


@pytest.mark.parametrize("nx,m", [(300, 100), (700, 324), (1500, 1001), (2200, 900)])
def test_bunch_kaufman_cluster_kkt(ctx, nx, m):
    from hiop_b200.engine import LinSolverSymDense
    K = synth.make_kkt_like(nx, m, seed=nx + m)
    ret, ok, rhs, xs = _factor_solve(ctx, K, LinSolverSymDense.BUNCH_KAUFMAN, nrhs=1, seed=m)
    assert ret == m and ok
    ref = np.linalg.solve(K, rhs.T).T
    assert np.abs(xs - ref).max() <= 1e-8 * np.abs(ref).max()


def test_bunch_kaufman_cluster_mixed_scales_and_singular(ctx):
    from hiop_b200.engine import LinSolverSymDense
    r = np.random.default_rng(11)
    N = 900
    A = r.standard_normal((N, N))
    M = A + A.T
    M[np.diag_indices(N)] *= r.choice([1e-3, 1.0, 10.0], N)
    ev = np.linalg.eigvalsh(M)
    ret, ok, rhs, xs = _factor_solve(ctx, M, LinSolverSymDense.BUNCH_KAUFMAN, seed=3)
    assert ret == int((ev < 0).sum()) and ok
    assert np.abs(M @ xs.T - rhs.T).max() <= 1e-8 * np.abs(rhs).max() * max(1.0, np.linalg.cond(M) * 1e-3)
    K = synth.make_kkt_like(500, 120, seed=3)
    K[3, :] = 0.0
    K[:, 3] = 0.0
    ret, ok, _, _ = _factor_solve(ctx, K, LinSolverSymDense.BUNCH_KAUFMAN)
    assert ret == -1
