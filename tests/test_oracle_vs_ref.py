"""Pins the numpy/C restatement (oracle/kkt_oracle.py) against the UNMODIFIED reference (oracle/_ref).
Runs only where oracle/_ref was built (this container); the same tuples are frozen in tests/golden for the GPU box."""
import numpy as np
import pytest

from hiop_b200 import synth
from oracle import kkt_oracle as ko
from oracle import ref

pytestmark = pytest.mark.ref


def _ref_system(p):
    q = ref.RefQn(p.n, p.m_eq, p.m_ineq, max(p.l, 1), p.ixl, p.ixu, p.idl, p.idu)
    q.set_iterate(p.sxl, p.sxu, p.zl, p.zu, p.sdl, p.sdu, p.vl, p.vu)
    q.set_jac(p.Jc, p.Jd)
    q.set_secant(p.sigma, p.St, p.Yt, p.L, p.D)
    return q


def _oracle_state(p):
    Dx, DhInv, Dd, Dd_inv = ko.kkt_update(p.zl, p.sxl, p.zu, p.sxu, p.ixl, p.ixu, p.vl, p.sdl, p.vu, p.sdu, p.idl,
                                          p.idu, p.sigma)
    return Dx, ko.QnState(p.Jc, p.Jd, DhInv, Dd_inv, p.St, p.Yt, p.L, p.D, p.sigma)


@pytest.mark.parametrize("n,m,l,mz", [(500, 4, 6, False), (2000, 50, 6, True), (1000, 1, 3, False),
                                       (3000, 64, 0, False), (777, 33, 2, True)])
def test_qn_system_matches_reference(n, m, l, mz):
    p = synth.make_qn_problem(n, m, l, masked_zero_divisors=mz)
    q = _ref_system(p)
    Dx_r, DhInv_r, Ddinv_r = q.update()
    Dx, st = _oracle_state(p)
    np.testing.assert_array_equal(Dx, Dx_r)
    np.testing.assert_array_equal(st.DhInv, DhInv_r)
    np.testing.assert_array_equal(st.Dd_inv, Ddinv_r)
    N_r = q.condense()
    N, W0, S1, Y1 = ko.condense(st)
    scale = np.abs(N_r).max()
    assert np.abs(N - N_r).max() <= 1e-12 * scale
    hs_r = q.hess_solve(p.rx)
    hs = ko.hess_solve(st, p.rx)
    assert np.abs(hs - hs_r).max() <= 1e-11 * np.abs(hs_r).max()
    dx_r, dyc_r, dyd_r = q.solve_compressed(p.rx, p.ryc, p.ryd)
    dx, dyc, dyd, _ = ko.solve_compressed(st, p.rx, p.ryc, p.ryd)
    for a, b in ((dx, dx_r), (dyc, dyc_r), (dyd, dyd_r)):
        if b.size:
            assert np.abs(a - b).max() <= 1e-8 * max(1.0, np.abs(b).max())
    q.close()


def test_compute_directions_matches_reference():
    p = synth.make_qn_problem(1500, 20, 4, masked_zero_divisors=True)
    q = _ref_system(p)
    q.update()
    _, st = _oracle_state(p)
    d_r = q.compute_directions(p.res)
    it = dict(sxl=p.sxl, sxu=p.sxu, zl=p.zl, zu=p.zu, sdl=p.sdl, sdu=p.sdu, vl=p.vl, vu=p.vu)
    pat = dict(ixl=p.ixl, ixu=p.ixu, idl=p.idl, idu=p.idu)
    d = ko.compute_directions(st, it, pat, p.res)
    for k in ko.DIR_NAMES:
        assert np.all(np.isfinite(d_r[k])), k
        assert np.abs(d[k] - d_r[k]).max() <= 1e-8 * max(1.0, np.abs(d_r[k]).max()), k
    q.close()


def test_hess_times_vec_matches_reference():
    p = synth.make_qn_problem(900, 3, 5)
    q = _ref_system(p)
    Dx_r, _, _ = q.update()
    x = np.random.default_rng(3).standard_normal(p.n)
    y0 = np.random.default_rng(4).standard_normal(p.n)
    for add in (False, True):
        y_r = q.hess_times_vec(0.5, y0, 2.0, x, add)
        y = ko.hess_times_vec(p.St, p.Yt, p.sigma, Dx_r, 0.5, y0, 2.0, x, add)
        assert np.abs(y - y_r).max() <= 1e-10 * np.abs(y_r).max()
    q.close()


@pytest.mark.parametrize("nx,m", [(30, 10), (120, 37), (5, 0), (1, 1)])
def test_symdense_matches_reference(nx, m):
    K = synth.make_kkt_like(nx, m)
    rhs = np.random.default_rng(5).standard_normal(nx + m)
    ret_r, sol_r, _, _ = ref.symdense_factor_solve(np.triu(K), rhs)
    ret, f = ko.symdense_matrix_changed(np.triu(K))
    assert ret == ret_r == m
    sol = f.solve(rhs)
    assert np.abs(sol - sol_r).max() <= 1e-9 * np.abs(sol_r).max()
    assert np.abs(K @ sol - rhs).max() <= 1e-9 * np.abs(rhs).max()


def test_symdense_singular_matches_reference():
    K = synth.make_kkt_like(20, 6)
    K[3, :] = 0.0
    K[:, 3] = 0.0
    ret_r, _, _, _ = ref.symdense_factor_solve(np.triu(K))
    ret, _ = ko.symdense_matrix_changed(np.triu(K))
    assert ret == ret_r == -1


def test_vector_ops_match_reference():
    r = np.random.default_rng(9)
    n = 1000
    y, x = r.standard_normal(n), r.standard_normal(n)
    z = r.uniform(0.5, 2.0, n)
    sel = (r.random(n) < 0.6).astype(np.float64)
    z0 = z * sel  # zero divisors on masked-out lanes
    for alpha in (1.0, -1.0, 0.37):
        yr, _ = ref.vec_op("axdzpy_w_pattern", y, x, z0, sel, alpha)
        np.testing.assert_array_equal(ko.axdzpy_w_pattern(y.copy(), alpha, x, z0, sel), yr)
        yr, _ = ref.vec_op("axzpy", y, x, z, None, alpha)
        np.testing.assert_array_equal(ko.axzpy(y.copy(), alpha, x, z), yr)
    yr, _ = ref.vec_op("component_div_w_sel", y, z0, None, sel)
    np.testing.assert_array_equal(ko.component_div_w_select(y.copy(), z0, sel), yr)
    yr, _ = ref.vec_op("add_logbar_grad", y, z0, None, sel, 0.1)
    np.testing.assert_array_equal(ko.add_log_barrier_grad(y.copy(), 0.1, z0, sel), yr)
    _, lb = ref.vec_op("logbarrier", z, None, None, sel)
    assert lb == ko.log_barrier(z, sel)
    ixu = (r.random(n) < 0.3).astype(np.float64)
    _, ld = ref.vec_op("lin_damping_term", z, sel, ixu, None, 0.1, 1e-5)
    assert ld == ko.linear_damping_term(z, sel, ixu, 0.1, 1e-5)
    yr, _ = ref.vec_op("add_lin_damping", y, sel, ixu, None, 0.9, 1e-6)
    np.testing.assert_array_equal(ko.add_linear_damping_term(y.copy(), sel, ixu, 0.9, 1e-6), yr)
    _, fb = ref.vec_op("frac_to_bdry_w_sel", z, x, None, sel, 0.995)
    assert fb == ko.fraction_to_the_bdry(z, x, 0.995, sel)
    _, fb = ref.vec_op("frac_to_bdry", z, x, None, None, 0.995)
    assert fb == ko.fraction_to_the_bdry(z, x, 0.995)


def test_mds_assembly_ops_match_reference():
    r = np.random.default_rng(21)
    Nw, m, n = 40, 7, 12
    A = r.standard_normal((m, n))
    W = r.standard_normal((Nw, Nw))
    Wr = W.copy()
    ref.lib().ref_mat_trans_add_to_sym_upper(m, n, A.ctypes.data_as(ref.dp), 3, 20, 0.7, Nw, Wr.ctypes.data_as(ref.dp))
    np.testing.assert_array_equal(ko.trans_add_to_sym_upper(A, 3, 20, 0.7, W.copy()), Wr)
    H = r.standard_normal((n, n))
    Wr = W.copy()
    ref.lib().ref_mat_add_upper_to_sym_upper(n, H.ctypes.data_as(ref.dp), 5, -1.3, Nw, Wr.ctypes.data_as(ref.dp))
    np.testing.assert_array_equal(ko.add_upper_to_sym_upper(H, 5, -1.3, W.copy()), Wr)
    # sparse Schur terms: sorted triplets, ~4 nnz per row
    ms, ns = 9, 30
    rows, cols = [], []
    for i in range(ms):
        cs = np.sort(r.choice(ns, 4, replace=False))
        rows += [i] * 4
        cols += list(cs)
    iR, jC = np.array(rows, dtype=np.int32), np.array(cols, dtype=np.int32)
    vals = r.standard_normal(iR.size)
    D = r.uniform(0.5, 2.0, ns)
    Wr = W.copy()
    ref.lib().ref_sp_add_MDinvMtrans(ms, ns, iR.size, iR.ctypes.data_as(ref.ip), jC.ctypes.data_as(ref.ip),
                                     vals.ctypes.data_as(ref.dp), 11, -1.0, D.ctypes.data_as(ref.dp), Nw,
                                     Wr.ctypes.data_as(ref.dp))
    Wo = ko.sp_add_MDinvMtrans(ms, ns, iR, jC, vals, 11, -1.0, D, W.copy())
    assert np.abs(Wo - Wr).max() <= 1e-13 * np.abs(Wr).max()
