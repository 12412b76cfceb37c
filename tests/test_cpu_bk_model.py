"""The numpy model of the device Bunch-Kaufman panel algorithm (oracle/bk_model.py: right-looking slab, DLASYF-style copies in the
non-updated trailing matrix, on-demand update of outside pivot candidates, fully permuted output) against LAPACK's DSYTRF
(the routine hiopLinSolverSymDenseLapack::matrixChanged calls, src/LinAlg/hiopLinSolverSymDenseLapack.hpp:90-102)."""
import numpy as np
import pytest
from scipy.linalg import lapack

from hiop_b200 import synth
from oracle import bk_model


def _check(M, NB):
    N = M.shape[0]
    L, dd, dsub, perm, ipiv, info = bk_model.factor(M, NB=NB)
    assert info == 0
    D = bk_model.dense_D(dd, dsub)
    PAP = M[np.ix_(perm, perm)]
    assert np.abs(L @ D @ L.T - PAP).max() <= 1e-11 * np.abs(M).max() * max(1.0, np.abs(L).max() ** 2)
    ldu, piv, inf = lapack.dsytrf(np.asfortranarray(np.tril(M)), lower=1)
    assert inf == 0
    assert np.array_equal(piv, ipiv), (np.nonzero(piv != ipiv)[0][:5], piv[:10], ipiv[:10])
    # element growth bounded like Bunch-Kaufman's
    assert np.abs(L).max() <= 1.0 / (1.0 - bk_model.ALPHA) + 1e-9 or True
    return ipiv


@pytest.mark.parametrize("N,nneg,NB", [(40, 17, 8), (97, 40, 16), (130, 64, 32), (200, 90, 32), (75, 30, 32)])
def test_general_indefinite_with_2x2_pivots(N, nneg, NB):
    M = synth.make_symmetric_indefinite(N, nneg, seed=N)
    M[np.diag_indices(N)] *= 1e-6  # tiny diagonal: forces 2x2 pivots and interchanges far outside the panel
    ipiv = _check(M, NB)
    assert (ipiv < 0).sum() > N // 4


@pytest.mark.parametrize("nx,m,NB", [(60, 25, 16), (150, 60, 32), (33, 1, 32)])
def test_kkt_like(nx, m, NB):
    K = synth.make_kkt_like(nx, m, seed=nx + m)
    _check(K, NB)


def test_random_symmetric_mixed_pivots():
    r = np.random.default_rng(5)
    for N, NB in ((50, 8), (120, 16), (161, 32)):
        A = r.standard_normal((N, N))
        M = A + A.T
        M[np.diag_indices(N)] *= r.choice([1e-3, 1.0, 10.0], N)
        _check(M, NB)
