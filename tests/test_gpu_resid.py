"""hiopResidual::update on the device against the oracle restatement (pinned to the reference by
tests/test_oracle_vs_ref.py::test_residual_update_matches_reference): elementwise blocks bit for bit, rx to 1e-13, the 11 norms;
then the residual feeds hb_lowrank_compute_directions without leaving the device."""
import numpy as np
import pytest

from hiop_b200 import synth
from oracle import kkt_oracle as ko
from test_gpu_parity import ctx, _setup_kkt, _as_dict  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,mz,mu,kd", [(20000, 40, True, 0.1, 1e-5), (4099, 37, False, 1e-4, 0.0), (3000, 1, True, 1.0, 1e-5), (2500, 0, False, 0.5, 1e-5)])
def test_residual_update_against_oracle(ctx, n, m, mz, mu, kd):
    P = synth.make_qn_problem(n, m, 0, masked_zero_divisors=mz, seed=17 + n)
    p = _as_dict(P)
    itr, dat = synth.make_iterate(P)
    k, T = _setup_kkt(ctx, p)
    D = ctx.to_device
    it_d = {kk: D(np.ascontiguousarray(v)) for kk, v in itr.items()}
    sizes = {kk: v.size for kk, v in itr.items()}
    res_d = {rk: ctx.zeros(sizes[dk]) for rk, dk in zip(ko.RES_NAMES, ko.DIR_NAMES)}
    nm = k.residual_update(it_d, D(dat["c"]), D(dat["d"]), D(dat["grad"]), mu, kd, D(dat["xl"]), D(dat["xu"]), D(dat["dl"]), D(dat["du"]), D(dat["crhs"]), res_d)
    ctx.sync()
    pat = dict(ixl=P.ixl, ixu=P.ixu, idl=P.idl, idu=P.idu)
    ro, no = ko.residual_update(itr, dat["c"], dat["d"], dat["grad"], P.Jc, P.Jd, mu, kd, pat, dat["xl"], dat["xu"], dat["dl"], dat["du"], dat["crhs"])
    for rk in ko.RES_NAMES:
        got = res_d[rk].cpu().numpy()
        if rk == "rx":
            assert np.abs(got - ro[rk]).max(initial=0.0) <= 1e-13 * max(1.0, np.abs(ro[rk]).max(initial=0.0))
        else:
            np.testing.assert_array_equal(got, ro[rk], err_msg=rk)
    for kk in ko.NORM_NAMES:
        assert abs(nm[kk] - no[kk]) <= 1e-12 * max(1.0, abs(no[kk])), (kk, nm[kk], no[kk])
    if m:
        # the residual blocks are exactly what compute_directions consumes: solve and compare with the oracle on the oracle's residual
        dirs = {kk: ctx.zeros(sizes[kk]) for kk in ko.DIR_NAMES}
        assert k.computeDirections(res_d, dirs)
        ctx.sync()
        Dx, DhInv, Dd, Dd_inv = ko.kkt_update(P.zl, P.sxl, P.zu, P.sxu, P.ixl, P.ixu, P.vl, P.sdl, P.vu, P.sdu, P.idl, P.idu, P.sigma)
        st = ko.QnState(P.Jc, P.Jd, DhInv, Dd_inv, P.St, P.Yt, P.L, P.D, P.sigma)
        itk = dict(sxl=P.sxl, sxu=P.sxu, zl=P.zl, zu=P.zu, sdl=P.sdl, sdu=P.sdu, vl=P.vl, vu=P.vu)
        do = ko.compute_directions(st, itk, pat, ro)
        for kk in ko.DIR_NAMES:
            b = do[kk]
            assert np.abs(dirs[kk].cpu().numpy() - b).max(initial=0.0) <= 1e-8 * max(1.0, np.abs(b).max(initial=0.0)), kk
    k.close()
