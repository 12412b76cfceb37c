"""The reference's own closed-form known-answer tests for the kernels on the hot path, replayed against the oracle
restatement with the reference's constants. Each test cites the KAT it restates (tests/LinAlg/vectorTests.hpp,
matrixTestsDense.hpp, matrixTestsSparse.hpp of the reference tree); sizes follow the reference's test drivers (small local sizes).
The same operations are pinned against the compiled reference on random data in tests/test_oracle_vs_ref.py; this file is the
"every known-answer test the reference holds for the path" part of the oracle's pinning."""
import numpy as np
import pytest

from oracle import kkt_oracle as ko

N = 100   # local_ordinal_type Nlocal in tests/testVector.cpp is 1000 per rank; the KATs are size-independent


def _c(v, n=N):
    return np.full(n, float(v))


def test_vectorAxdzpy_w_patternSelect():
    # vectorTests.hpp:1167-1203: v = 2, x = 1/2, z = 1/2, alpha = 3; last pattern entry 0 with z = 0 there
    v, x, z, p = _c(2), _c(0.5), _c(0.5), _c(1)
    p[-1] = 0.0
    z[-1] = 0.0
    out = ko.axdzpy_w_pattern(v.copy(), 3.0, x, z, p)
    want = _c(2.0 + 3.0 * 0.5 / 0.5)
    want[-1] = 2.0
    np.testing.assert_array_equal(out, want)


def test_vectorComponentDiv_p_selectPattern():
    # vectorTests.hpp:833-866: v = 1/2, x = 1; masked-out entry (x = 0) becomes ZERO, not kept
    v, x, p = _c(0.5), _c(1), _c(1)
    p[-1] = 0.0
    x[-1] = 0.0
    out = ko.component_div_w_select(v.copy(), x, p)
    want = _c(0.5)
    want[-1] = 0.0
    np.testing.assert_array_equal(out, want)


def test_vectorInvert():
    # vectorTests.hpp:1296-1303
    np.testing.assert_array_equal(1.0 / _c(2), _c(0.5))


def test_vectorLogBarrier():
    # vectorTests.hpp:1309-1343: log(1) over N-1 selected entries (the unselected one holds 3000); then a single selected entry
    x, p = _c(1), _c(1)
    p[-1] = 0.0
    x[-1] = 3000.0
    assert ko.log_barrier(x, p) == (N - 1) * np.log(1.0)
    x, p = _c(0), _c(0)
    p[-1] = 1.0
    x[-1] = 1.0
    assert ko.log_barrier(x, p) == np.log(1.0)


def test_vectorAddLogBarrierGrad():
    # vectorTests.hpp:1372-1403: x = 2, y = 2, alpha = 1/2 -> x + alpha / y on the pattern
    x, y, p = _c(2), _c(2), _c(1)
    p[-1] = 0.0
    out = ko.add_log_barrier_grad(x.copy(), 0.5, y, p)
    want = _c(2.0 + 0.5 / 2.0)
    want[-1] = 2.0
    np.testing.assert_array_equal(out, want)


def test_vectorLinearDampingTerm():
    # vectorTests.hpp:1417-1453: x = 1, left = 1, right = 0 except one entry with left = right = 2 (counts only left==1 & right==0)
    x, left, right = _c(1), _c(1), _c(0)
    left[-1] = 2.0
    right[-1] = 2.0
    expected = float(N - 1)
    expected *= 2.0
    expected *= 2.0
    assert ko.linear_damping_term(x, left, right, 2.0, 2.0) == expected


def test_vectorAddLinearDampingTerm():
    # vectorTests.hpp:1461-1540: x = alpha*x + ct*(left - right) with the four (left, right) combinations, alpha = 1/4, ct = 2
    x, left, right = _c(1), _c(1), _c(0)
    left[0], right[0] = 0.0, 1.0
    left[1], right[1] = 0.0, 0.0
    left[2], right[2] = 1.0, 1.0
    left[3], right[3] = 1.0, 0.0
    out = ko.add_linear_damping_term(x.copy(), left, right, 0.25, 2.0)
    assert out[0] == 1.0 * 0.25 - 2.0
    assert out[1] == 1.0 * 0.25
    assert out[2] == 1.0 * 0.25
    assert out[3] == 1.0 * 0.25 + 2.0
    assert np.all(out[4:] == 0.25 + 2.0)


def test_vectorMin_w_pattern():
    # vectorTests.hpp:1630-1650 (the restated reduction is the masked minimum used by hiopResidual::update / adjust_small_slacks)
    x, p = _c(1), _c(1)
    x[-1] = -1.0
    assert min(1e100, x[p == 1.0].min()) == -1.0
    p[-1] = 0.0
    assert min(1e100, x[p == 1.0].min()) == 1.0


def test_vectorFractionToTheBdry_and_w_pattern():
    # vectorTests.hpp:1773-1850: tau = 1/2, x = 1
    x = _c(1)
    assert ko.fraction_to_the_bdry(x, _c(2), 0.5) == 1.0                    # default when dx >= 0
    dx = _c(-1)
    dx[-1] = -2.0
    assert ko.fraction_to_the_bdry(x, dx, 0.5) == 0.25                      # -0.5 * 1 / (-2)
    p = _c(1)
    assert ko.fraction_to_the_bdry(x, _c(1), 0.5, p) == 1.0
    p[-1] = 0.0
    dx = _c(1)
    dx[-1] = -0.5
    assert ko.fraction_to_the_bdry(x, dx, 0.5, p) == 1.0                    # the only negative step is masked out
    p = _c(1)
    dx = _c(-1)
    dx[-1] = -2.0
    assert ko.fraction_to_the_bdry(x, dx, 0.5, p) == 0.25


def test_vectorAdjustDuals_plh():
    # vectorTests.hpp:1891-1940: z = 1, x = 2, mu = kappa = 1/2 -> a = mu/x = 1/4, b = a/kappa = 1/2, a*kappa = 1/8: z >= b and
    # a <= b -> z = b
    out = ko.adjust_duals_plh(_c(1), _c(2), _c(1), 0.5, 0.5)
    np.testing.assert_array_equal(out, _c(0.5))


def test_matrixTimesVec_and_TransTimesVec():
    # matrixTestsDense.hpp:173-200, 208-258: A = 1, x = y = 3, alpha = beta = 1
    M, Nc = 7, 12
    A = np.ones((M, Nc))
    y = np.full(M, 3.0)
    ko.times_vec(A, 1.0, y, 1.0, np.full(Nc, 3.0))
    np.testing.assert_array_equal(y, np.full(M, 3.0 + 1.0 * 3.0 * Nc))
    A = np.ones((M, Nc))
    A[:, Nc - 1] = 0.0                                                     # zero a row of A^T
    y = np.full(Nc, 3.0)
    ko.trans_times_vec(A, 1.0, y, 1.0, np.full(M, 3.0))
    want = np.full(Nc, 3.0 + 3.0 * M)
    want[Nc - 1] = 3.0
    np.testing.assert_array_equal(y, want)


def test_matrixAddSubDiagonal():
    # matrixTestsDense.hpp:447-470: A = 1/2, x = 1, alpha = 1/2, the vector lands at the END of the diagonal
    Nn, xl = 9, 4
    A = np.full((Nn, Nn), 0.5)
    ko.add_sub_diagonal(A, Nn - xl, 0.5, np.ones(xl))
    want = np.full((Nn, Nn), 0.5)
    for i in range(Nn - xl, Nn):
        want[i, i] = 0.5 + 1.0 * 0.5
    np.testing.assert_array_equal(A, want)


def test_matrixTransAddToSymDenseMatrixUpperTriangle():
    # matrixTestsDense.hpp:544-575: W = 1, A = 1/2 (A_M x A_N), alpha = 1/2; A^T lands at rows [0, A_N), columns [N - A_M, N)
    Nw, AM, AN = 10, 3, 5
    W = np.ones((Nw, Nw))
    ko.trans_add_to_sym_upper(np.full((AM, AN), 0.5), 0, Nw - AM, 0.5, W)
    want = np.ones((Nw, Nw))
    want[0:AN, Nw - AM:Nw] = 1.0 + 0.5 * 0.5
    np.testing.assert_array_equal(W, want)


def test_matrixAddUpperTriangleToSymDenseMatrixUpperTriangle():
    # matrixTestsDense.hpp:587-620: only the upper triangle of A (incl. diagonal) is added, at W's upper-left corner
    Nw, An = 10, 4
    W = np.ones((Nw, Nw))
    ko.add_upper_to_sym_upper(np.full((An, An), 0.5), 0, 0.5, W)
    want = np.ones((Nw, Nw))
    for i in range(An):
        for j in range(i, An):
            want[i, j] = 1.0 + 0.5 * 0.5
    np.testing.assert_array_equal(W, want)


def _sparse_pattern(m, n, per_row, seed):
    rng = np.random.default_rng(seed)
    iRow, jCol = [], []
    for i in range(m):
        cols = np.sort(rng.choice(n, size=per_row, replace=False))
        iRow += [i] * per_row
        jCol += list(cols)
    return np.array(iRow, dtype=np.int32), np.array(jCol, dtype=np.int32)


@pytest.mark.parametrize("offset", [0, 3])
def test_matrixAddMDinvMtransToDiagBlockOfSymDeMatUTri(offset):
    # matrixTestsSparse.hpp:415-490: A = 1 on its pattern, D = 1/2, W = 0, alpha = 1/2: W_ij (upper, inside the block) =
    # alpha * (#columns shared by rows i and j) / d
    m, n = 6, 15
    iRow, jCol = _sparse_pattern(m, n, 4, 1)
    W = np.zeros((m + offset + 2, m + offset + 2))
    ko.sp_add_MDinvMtrans(m, n, iRow, jCol, np.ones(iRow.size), offset, 0.5, np.full(n, 0.5), W)
    rows = [set(jCol[iRow == i]) for i in range(m)]
    want = np.zeros_like(W)
    for i in range(m):
        for j in range(i, m):
            want[offset + i, offset + j] = 0.5 * len(rows[i] & rows[j]) * 1.0 * 1.0 / 0.5
    np.testing.assert_array_equal(W, want)


def test_matrixAddMDinvNtransToSymDeMatUTri():
    # matrixTestsSparse.hpp:587-680: same with two different sparse matrices (rows of M against rows of N), block placed at
    # (row_start, col_start) of the upper triangle
    m1, m2, n = 5, 4, 12
    iR1, jC1 = _sparse_pattern(m1, n, 3, 2)
    iR2, jC2 = _sparse_pattern(m2, n, 5, 3)
    W = np.zeros((m1 + m2 + 1, m1 + m2 + 1))
    ko.sp_add_MDinvNtrans(m1, n, iR1, jC1, np.ones(iR1.size), m2, iR2, jC2, np.ones(iR2.size), 0, m1, 0.5, np.full(n, 0.5), W)
    r1 = [set(jC1[iR1 == i]) for i in range(m1)]
    r2 = [set(jC2[iR2 == i]) for i in range(m2)]
    want = np.zeros_like(W)
    for i in range(m1):
        for j in range(m2):
            want[i, m1 + j] = 0.5 * len(r1[i] & r2[j]) / 0.5
    np.testing.assert_array_equal(W, want)
