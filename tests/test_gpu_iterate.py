"""hiopIterate / hiopLogBarProblem vector pipeline on the device: fraction-to-the-boundary, step, log-barrier function
and gradients, against the oracle restatement (pinned to the reference by
tests/test_oracle_vs_ref.py::test_logbar_and_fraction_to_bdry_match_reference)."""
import numpy as np
import pytest

from hiop_b200 import synth
from oracle import kkt_oracle as ko
from test_gpu_parity import ctx, _setup_kkt, _as_dict  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,mz,mu,kd", [(20000, 40, True, 0.1, 1e-5), (4099, 37, False, 1e-4, 0.0), (3000, 1, True, 1.0, 1e-5), (2500, 0, False, 0.5, 1e-5)])
def test_line_search_pipeline_against_oracle(ctx, n, m, mz, mu, kd):
    P = synth.make_qn_problem(n, m, 0, masked_zero_divisors=mz, seed=23 + n)
    p = _as_dict(P)
    itr, dat = synth.make_iterate(P)
    pat = dict(ixl=P.ixl, ixu=P.ixu, idl=P.idl, idu=P.idu)
    for s, ptn in (("sxl", "ixl"), ("sxu", "ixu"), ("sdl", "idl"), ("sdu", "idu")):
        itr[s] = np.where(pat[ptn] == 1.0, np.abs(itr[s]) + 1e-3, itr[s])
    rng = np.random.default_rng(5)
    direction = {kk: rng.standard_normal(np.asarray(v).size) * np.where(np.asarray(v) != 0, 1.0, 0.0) for kk, v in itr.items()}
    k, T = _setup_kkt(ctx, p)
    D = ctx.to_device
    it_d = {kk: D(np.ascontiguousarray(v)) for kk, v in itr.items()}
    dir_d = {kk: D(np.ascontiguousarray(v)) for kk, v in direction.items()}
    # fraction to the boundary: exact (a minimum of identically computed quotients)
    ap, ad = k.fraction_to_bdry(it_d, dir_d, 0.995)
    apo, ado = ko.iterate_fraction_to_bdry(itr, direction, 0.995, pat)
    assert (ap, ad) == (apo, ado)
    # trial point
    out_d = {kk: ctx.zeros(np.asarray(v).size) for kk, v in itr.items()}
    k.take_step(it_d, dir_d, ap, ad, out_d)
    ctx.sync()
    for kk in ko.DIR_NAMES:
        al = ap if kk in ("x", "d", "yc", "yd") else ad
        if kk in ("sxl", "sxu", "sdl", "sdu"):
            continue                                 # slacks are recomputed from x, d by the driver (hiopIterate.cpp:270-303)
        want = itr[kk] + al * direction[kk]
        assert np.abs(out_d[kk].cpu().numpy() - want).max(initial=0.0) <= 1e-15 * max(1.0, np.abs(want).max(initial=0.0)), kk
    # log-barrier function + gradients, and the function-only variant
    gx, gd = ctx.zeros(n), ctx.zeros(P.m_ineq)
    fl = k.logbar(it_d, 3.25, mu, kd, D(dat["grad"]), gx, gd)
    flo, gxo, gdo = ko.logbar_update(itr, 3.25, mu, kd, dat["grad"], pat)
    ctx.sync()
    assert abs(fl - flo) <= 1e-12 * max(1.0, abs(flo))
    np.testing.assert_array_equal(gx.cpu().numpy(), gxo)
    np.testing.assert_array_equal(gd.cpu().numpy(), gdo)
    assert k.logbar(it_d, 3.25, mu, kd) == fl
    # adjustDuals_primalLogHessian: every branch of the clamp (duals spread over 12 decades), in place, bit for bit
    itr2 = dict(itr)
    for zk in ("zl", "zu", "vl", "vu"):
        itr2[zk] = itr[zk] * 10.0 ** rng.integers(-6, 7, size=np.asarray(itr[zk]).size)
    it2_d = {kk: D(np.ascontiguousarray(v)) for kk, v in itr2.items()}
    k.adjust_duals_plh(it2_d, mu, 50.0)
    ctx.sync()
    for name, want in zip(("zl", "zu", "vl", "vu"), ko.iterate_adjust_duals(itr2, pat, mu, 50.0)):
        np.testing.assert_array_equal(it2_d[name].cpu().numpy(), want, err_msg=name)
    k.close()


@pytest.mark.parametrize("mu", [1e-2, 10.0])
def test_adjust_small_slacks_against_oracle(ctx, mu):
    """hiopIterate::adjust_small_slacks: collapsed slacks are pushed back exactly like the reference's chain of vector operations does
    (oracle pinned bit-for-bit in tests/test_oracle_vs_ref.py); blocks without a small slack stay untouched."""
    P = synth.make_qn_problem(5000, 24, 0, masked_zero_divisors=True, seed=3)
    p = _as_dict(P)
    itr, dat = synth.make_iterate(P)
    pat = dict(ixl=P.ixl, ixu=P.ixu, idl=P.idl, idu=P.idu)
    rng = np.random.default_rng(4)
    k, T = _setup_kkt(ctx, p)
    D = ctx.to_device
    cur_d = {kk: D(np.ascontiguousarray(v)) for kk, v in itr.items()}
    bounds = [D(dat[b]) for b in ("xl", "xu", "dl", "du")]
    for collapse in (True, False):
        trial = {kk: np.array(v, dtype=np.float64) for kk, v in itr.items()}
        for s_, ptn in (("sxl", "ixl"), ("sxu", "ixu"), ("sdl", "idl"), ("sdu", "idu")):
            trial[s_] = np.where(pat[ptn] == 1.0, np.abs(trial[s_]) + 1e-3, 0.0)
            if collapse:
                hit = (rng.random(trial[s_].size) < 0.2) & (pat[ptn] == 1.0)
                trial[s_] = np.where(hit, rng.choice([0.0, -1e-9, 1e-20, 3e-17], size=trial[s_].size), trial[s_])
        tr_d = {kk: D(np.ascontiguousarray(v)) for kk, v in trial.items()}
        num = k.adjust_small_slacks(tr_d, cur_d, mu, *bounds)
        ctx.sync()
        want_num = 0
        for s_, ptn, bnd, dual in (("sxl", "ixl", "xl", "zl"), ("sxu", "ixu", "xu", "zu"), ("sdl", "idl", "dl", "vl"), ("sdu", "idu", "du", "vu")):
            new, cnt = ko.adjust_small_slack(trial[s_], dat[bnd], itr[dual], pat[ptn], mu)
            want_num += cnt
            np.testing.assert_array_equal(tr_d[s_].cpu().numpy(), new, err_msg=s_)
        assert num == want_num and (num > 0) == collapse
    k.close()
