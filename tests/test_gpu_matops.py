"""GPU known-answer tests for the standalone hiopMatrixDenseRowMajor primitives (hb_matops.cu) -- the reference's own closed-form KATs
(tests/LinAlg/matrixTestsDense.hpp, constants and shapes as cited per test; tests/test_cpu_reference_kats.py replays the same KATs against
the oracle) plus seeded comparisons with the oracle's restatements, and hiopKKTLinSysCompressed::test_direction against its numpy
restatement (src/Optimization/hiopKKTLinSys.cpp:455-509)."""
import ctypes

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


def _p(t):
    # the pointer object keeps its tensor alive: `_p(ctx.to_device(a))` as a call argument would otherwise free the block before the
    # next argument is built, and torch's caching allocator may hand the same bytes out again
    q = ctypes.c_void_p(t.data_ptr())
    q._keep = t
    return q


def test_matrixTimesMatTrans(ctx):
    # matrixTestsDense.hpp:360-398: A = 2 (m x n), W = 2 (m x k), X = 3 (k x n) with its last row zeroed, alpha = beta = 2:
    # W(i,j) = beta*W + alpha*A*X*n, and beta*W in the column of the zero row
    L = ctx.L
    M, K, N = 7, 5, 300
    A = ctx.to_device(np.full((M, N), 2.0))
    Xh = np.full((K, N), 3.0)
    Xh[K - 1, :] = 0.0
    W = ctx.to_device(np.full((M, K), 2.0))
    assert L.hb_mat_times_mat_trans(ctx.h, M, K, N, _p(A), N, _p(ctx.to_device(Xh)), N, 2.0, _p(W), K, 2.0) == 0
    ctx.sync()
    want = np.full((M, K), 2.0 * 2.0 + 2.0 * 2.0 * 3.0 * N)
    want[:, K - 1] = 2.0 * 2.0
    np.testing.assert_array_equal(W.cpu().numpy(), want)
    # seeded, against numpy (warp-shuffle summation order: 1e-13 relative)
    r = np.random.default_rng(3)
    An, Bn, Wn = r.standard_normal((13, 1001)), r.standard_normal((6, 1001)), r.standard_normal((13, 6))
    Wd = ctx.to_device(Wn.copy())
    assert L.hb_mat_times_mat_trans(ctx.h, 13, 6, 1001, _p(ctx.to_device(An)), 1001, _p(ctx.to_device(Bn)), 1001, -1.0, _p(Wd), 6, 0.37) == 0
    ctx.sync()
    ref = -Wn + 0.37 * An @ Bn.T
    assert np.abs(Wd.cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()


def test_matrixAddSubDiagonal_and_AddDiagonal(ctx):
    # matrixTestsDense.hpp:447-470: A = 1/2, x = 1, alpha = 1/2, the vector lands at the END of the diagonal
    L = ctx.L
    Nn, xl = 9, 4
    A = ctx.to_device(np.full((Nn, Nn), 0.5))
    x = ctx.to_device(np.ones(xl))
    assert L.hb_mat_add_sub_diagonal(ctx.h, _p(A), Nn, Nn - xl, xl, 0.5, _p(x), 0) == 0
    ctx.sync()
    want = ko.add_sub_diagonal(np.full((Nn, Nn), 0.5), Nn - xl, 0.5, np.ones(xl))
    np.testing.assert_array_equal(A.cpu().numpy(), want)
    # the (dest start, source start, count) overload (:741-758) and the constant overload (:760-768); addDiagonal = start 0, all n
    A = ctx.to_device(np.zeros((Nn, Nn)))
    d = ctx.to_device(np.arange(1.0, 8.0))
    assert L.hb_mat_add_sub_diagonal(ctx.h, _p(A), Nn, 2, 3, 2.0, _p(d), 4) == 0     # M[2+i][2+i] += 2*d[4+i]
    assert L.hb_mat_add_sub_diagonal(ctx.h, _p(A), Nn, 6, 3, -1.5, None, 0) == 0      # constant on the last three
    ctx.sync()
    want = np.zeros((Nn, Nn))
    for i in range(3):
        want[2 + i, 2 + i] += 2.0 * (5.0 + i)
        want[6 + i, 6 + i] += -1.5
    np.testing.assert_array_equal(A.cpu().numpy(), want)


def test_matrixAddMatrix_copyRowsFrom_copyBlock(ctx):
    L = ctx.L
    r = np.random.default_rng(5)
    # addMatrix (matrixTestsDense.hpp:472-493 uses A = 1/2, B = 1, alpha = 1/2)
    A = ctx.to_device(np.full((6, 11), 0.5))
    B = ctx.to_device(np.ones((6, 11)))
    assert L.hb_mat_add_matrix(ctx.h, 6, 11, _p(A), 11, 0.5, _p(B), 11) == 0
    ctx.sync()
    np.testing.assert_array_equal(A.cpu().numpy(), np.full((6, 11), 1.0))
    # copyRowsFrom with an index list (matrixTestsDense.hpp:868-905): dst row i = src row idx[i]
    src = r.standard_normal((9, 14))
    idx = np.array([7, 0, 3, 3, 8], dtype=np.int32)
    import torch
    idx_d = torch.from_numpy(idx).to(ctx.device)
    dst = ctx.zeros(5 * 14).reshape(5, 14)
    assert L.hb_mat_copy_rows_from(ctx.h, 5, 14, _p(dst), 14, _p(ctx.to_device(src)), 14, _p(idx_d)) == 0
    ctx.sync()
    np.testing.assert_array_equal(dst.cpu().numpy(), src[idx])
    # copyBlockFromMatrix / copyFromMatrixBlock (matrixTestsDense.hpp:907-977): a block lands at / is taken from an offset
    big = ctx.to_device(np.zeros((10, 12)))
    blk = r.standard_normal((4, 5))
    assert L.hb_mat_copy_block(ctx.h, 4, 5, _p(big), 12, 3, 6, _p(ctx.to_device(blk)), 5, 0, 0) == 0
    out = ctx.zeros(2 * 3).reshape(2, 3)
    assert L.hb_mat_copy_block(ctx.h, 2, 3, _p(out), 3, 0, 0, _p(big), 12, 4, 7) == 0
    ctx.sync()
    want = np.zeros((10, 12))
    want[3:7, 6:11] = blk
    np.testing.assert_array_equal(big.cpu().numpy(), want)
    np.testing.assert_array_equal(out.cpu().numpy(), want[4:6, 7:10])


def test_matrixTransAddToSymDenseMatrixUpperTriangle_and_AddUpperTriangle(ctx):
    L = ctx.L
    # matrixTestsDense.hpp:544-575: W = 1, A = 1/2 (A_M x A_N), alpha = 1/2; A^T lands at rows [0, A_N), columns [N - A_M, N)
    Nw, AM, AN = 10, 3, 5
    W = ctx.to_device(np.ones((Nw, Nw)))
    assert L.hb_mat_trans_add_to_sym_upper(ctx.h, AM, AN, _p(ctx.to_device(np.full((AM, AN), 0.5))), AN, 0, Nw - AM, 0.5, _p(W), Nw) == 0
    ctx.sync()
    np.testing.assert_array_equal(W.cpu().numpy(), ko.trans_add_to_sym_upper(np.full((AM, AN), 0.5), 0, Nw - AM, 0.5, np.ones((Nw, Nw))))
    # matrixTestsDense.hpp:587-620: only the upper triangle of A (incl. diagonal) is added, at W's upper-left corner
    An = 4
    W = ctx.to_device(np.ones((Nw, Nw)))
    assert L.hb_mat_add_upper_to_sym_upper(ctx.h, An, _p(ctx.to_device(np.full((An, An), 0.5))), An, 0, 0.5, _p(W), Nw) == 0
    ctx.sync()
    np.testing.assert_array_equal(W.cpu().numpy(), ko.add_upper_to_sym_upper(np.full((An, An), 0.5), 0, 0.5, np.ones((Nw, Nw))))
    # seeded, offset block
    r = np.random.default_rng(9)
    A = r.standard_normal((6, 4))
    W0 = r.standard_normal((15, 15))
    W = ctx.to_device(W0.copy())
    assert L.hb_mat_trans_add_to_sym_upper(ctx.h, 6, 4, _p(ctx.to_device(A)), 4, 2, 8, -0.7, _p(W), 15) == 0
    ctx.sync()
    np.testing.assert_array_equal(W.cpu().numpy(), ko.trans_add_to_sym_upper(A, 2, 8, -0.7, W0.copy()))


def test_lowrank_test_direction(ctx):
    """dWd = dx^T (B + Dx + delta_wx) dx + dd^T (Dd + delta_wd) dd vs neg_curv_test_fact (||dx||^2 + ||dd||^2), B in the compact form that
    tests/test_gpu_parity.py pins against the reference's recursive timesVec (ko.hess_times_vec)."""
    from hiop_b200.engine import KKTLinSysLowRank
    P = synth.make_qn_problem(5000, 30, 5, seed=21)
    k = KKTLinSysLowRank(ctx, P.n, P.m_eq, P.m_ineq, 5)
    D = ctx.to_device
    J = D(P.J)
    T = {name: D(getattr(P, name)) for name in ("ixl", "ixu", "idl", "idu", "zl", "sxl", "zu", "sxu", "vl", "sdl", "vu", "sdu", "St", "Yt")}
    k.set_patterns(T["ixl"], T["ixu"], T["idl"], T["idu"])
    k.set_jacobian(J[:P.m_eq], J[P.m_eq:])
    k.set_secant(P.sigma, T["St"], T["Yt"], P.L, P.D)
    k.update(T["zl"], T["sxl"], T["zu"], T["sxu"], T["vl"], T["sdl"], T["vu"], T["sdu"])
    r = np.random.default_rng(2)
    dx, dd = r.standard_normal(P.n), r.standard_normal(P.m_ineq)
    dwx, dwd = np.full(P.n, 1e-3), np.full(P.m_ineq, 2e-3)
    Dx, DhInv, Dd, Dd_inv = ko.kkt_update(P.zl, P.sxl, P.zu, P.sxu, P.ixl, P.ixu, P.vl, P.sdl, P.vu, P.sdu, P.idl, P.idu, P.sigma)
    Bdx = ko.hess_times_vec(P.St, P.Yt, P.sigma, Dx, 0.0, np.zeros(P.n), 1.0, dx, False)
    dWd = Bdx @ dx + ((Dx + dwx) * dx) @ dx + ((Dd + dwd) * dd) @ dd
    xs = dx @ dx + dd @ dd
    out = (ctypes.c_double * 2)()
    dx_d, dd_d, dwx_d, dwd_d = D(dx), D(dd), D(dwx), D(dwd)   # kept alive: a pointer into a freed torch block may be handed out again
    for fact, deltas in ((1e-11, True), (1e30, True), (1e-11, False)):
        rc = ctx.L.hb_lowrank_test_direction(k.h, _p(dx_d), _p(dd_d), _p(dwx_d) if deltas else None, _p(dwd_d) if deltas else None, fact, out)
        want = dWd if deltas else dWd - (dwx * dx) @ dx - (dwd * dd) @ dd
        assert rc == (0 if want < xs * fact else 1)
        # the compact form of B against the recursive one: 1e-10 (tests/test_gpu_parity.py), relative to the size of the three terms
        scale = abs(Bdx @ dx) + ((Dx + dwx) * dx) @ dx + ((Dd + dwd) * dd) @ dd
        assert abs(out[0] - want) <= 1e-9 * scale, (out[0], want, scale)
        assert abs(out[1] - xs) <= 1e-12 * xs, (out[1], xs)
    k.close()
