"""Exactness claims of the int8-slice condensation, checked on a numpy integer model of the device path
(oracle/oz_model.py): digit range, exact reconstruction of the operands, int32 accumulator bound per K chunk, and the error of the
truncated product against the FP64 Gram matrix for 6 / 7 / 8 slices."""
import numpy as np
import pytest

from oracle import oz_model as oz


def _B(M, K, seed, decades=3.0):
    r = np.random.default_rng(seed)
    return r.standard_normal((M, K)) * 10.0 ** r.uniform(-decades, decades, size=(M, 1)) * np.sqrt(10.0 ** r.uniform(-3, 3, size=(1, K)))


@pytest.mark.parametrize("S", [6, 7, 8])
def test_digits_are_int8_and_reconstruct_the_operand(S):
    B = _B(9, 4000, 1)
    B[3] = 0.0                                               # an all-zero row (exponent 0, all digits 0)
    B[5, 7] = np.abs(B[5]).max() * 4                         # the row maximum itself
    e = oz.row_exponents(B)
    Q = oz.slices(B, e, S)
    assert np.abs(Q).max() <= 64                             # |q| <= 64 (the MMA operands are int8)
    assert np.abs(Q[1:]).max() <= 64 and Q[1:].min() >= -64
    rec = sum(Q[p].astype(np.float64) * 2.0 ** (-(6 + 7 * p)) for p in range(S))
    b = np.ldexp(B, -e[:, None])
    grid = 2.0 ** (-(6 + 7 * (S - 1)))
    assert np.abs(rec - b).max() <= 0.5 * grid               # rounded to the last slice's grid ...
    np.testing.assert_array_equal(rec, np.round(b / grid) * grid)   # ... exactly (round-half-even of the magic-number add)
    assert np.all(Q[:, 3, :] == 0)


def test_int32_accumulator_bound_per_chunk():
    # worst case operands: every digit at its extreme -> (t+1) * Kc * 2^12 must stay below 2^31 with Kc = 2^19 / S
    # (hb_ozaki.cu takes chunks of 128-column stages and drops one stage when the product would reach 2^31 exactly: S = 8)
    for S in (6, 7, 8):
        stages = (524288 // S) // 128
        while S * stages * 128 * 4096 >= 2 ** 31:
            stages -= 1
        assert S * stages * 128 * 64 * 64 < 2 ** 31
        assert stages >= (524288 // S) // 128 - 1
    B = _B(4, 3000, 2)
    C, info = oz.gram(B, 8, chunk_cols=700)                  # several chunks: the result must not depend on the chunking
    C1, _ = oz.gram(B, 8, chunk_cols=3000)
    assert info["max_abs_int32_accumulator"] < 2 ** 31
    assert np.abs(C - C1).max() <= 1e-15 * np.abs(C1).max()


@pytest.mark.parametrize("S,tol", [(6, 2e-10), (7, 2e-12), (8, 5e-14)])
def test_truncated_product_error_against_fp64(S, tol):
    # the tolerances are the ones tests/test_gpu_ozaki.py holds the device path to
    B = _B(24, 20000, 3, decades=2.0)
    C, _ = oz.gram(B, S)
    ref = B @ B.T
    assert np.abs(C - ref).max() <= tol * np.abs(ref).max()
    assert np.array_equal(C, C.T)                            # symmetric by construction (same integer products both ways)


def test_fused_sweep_identity_for_step_2_of_solveCompressed():
    """The identity behind k_oz_rowmax_dot: with t = [J; S; Y] (DhInv .* rx),
    J (H+Dx)^-1 rx = t_J - Z [sigma t_S; t_Y],  Z = U V^-1,  U = [sigma J DhInv S^T, J DhInv Y^T]
    -- checked against the oracle's own hess_solve + J product (hiopKKTLinSys.cpp:1146-1157, hiopHessianLowRank.cpp:495-540)."""
    from hiop_b200 import synth
    from oracle import kkt_oracle as ko
    for n, m, l in ((3000, 25, 5), (1200, 40, 1), (900, 10, 0)):
        P = synth.make_qn_problem(n, m, l, seed=3 + l)
        Dx, DhInv, Dd, Dd_inv = ko.kkt_update(P.zl, P.sxl, P.zu, P.sxu, P.ixl, P.ixu, P.vl, P.sdl, P.vu, P.sdu, P.idl, P.idu, P.sigma)
        st = ko.QnState(P.Jc, P.Jd, DhInv, Dd_inv, P.St, P.Yt, P.L, P.D, P.sigma)
        want = st.J @ ko.hess_solve(st, P.rx)
        w = DhInv * P.rx
        t_J = st.J @ w
        got = t_J
        if l:
            _, _, S1, Y1 = ko.condense(st)
            U = np.hstack([S1, Y1])                              # S1 already carries sigma
            Z = st.Vfac.solve(U.T.copy()).T                      # m x 2l
            p = np.concatenate([P.sigma * (P.St @ w), P.Yt @ w])
            got = t_J - Z @ p
        assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max())
