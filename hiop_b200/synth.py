"""Synthetic quasi-Newton KKT inputs (host, numpy) -- the generalisation of NlpDenseConsEx2 to arbitrary m that
BASELINE.json's configs need (the bundled drivers have m <= 4).

Distributions follow SURVEY.md section 8(d): J rows i.i.d. N(0,1)/sqrt(n) plus one dense all-ones row (Ex2's
constraints are sums over all variables, src/Drivers/Dense/NlpDenseConsEx2.cpp:226-298), slacks/duals U(1e-3,1),
lower bounds everywhere and upper bounds on 10% of the entries (Ex2 pattern, NlpDenseConsEx2.cpp:53-81),
S ~ N(0,1), Y = S*U(0.5,2) so that s'y > 0, and L, D exactly as hiopHessianLowRank::update builds them
(src/Optimization/hiopHessianLowRank.cpp:262-388, growL/growD :779-821).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class QnProblem:
    n: int
    m_eq: int
    m_ineq: int
    l: int
    sigma: float
    Jc: np.ndarray
    Jd: np.ndarray
    ixl: np.ndarray
    ixu: np.ndarray
    idl: np.ndarray
    idu: np.ndarray
    sxl: np.ndarray
    sxu: np.ndarray
    zl: np.ndarray
    zu: np.ndarray
    sdl: np.ndarray
    sdu: np.ndarray
    vl: np.ndarray
    vu: np.ndarray
    St: np.ndarray
    Yt: np.ndarray
    L: np.ndarray
    D: np.ndarray
    rx: np.ndarray
    ryc: np.ndarray
    ryd: np.ndarray
    res: dict = field(default_factory=dict)   # the 12 residual blocks for computeDirections

    @property
    def m(self):
        return self.m_eq + self.m_ineq

    @property
    def J(self):
        return np.ascontiguousarray(np.vstack([self.Jc, self.Jd]))


def secant_LD(St: np.ndarray, Yt: np.ndarray):
    """L_ij = s_i^T y_j for i > j (else 0), D_i = s_i^T y_i."""
    SY = St @ Yt.T
    return np.tril(SY, -1).copy(), np.diag(SY).copy()


def make_qn_problem(n: int, m: int, l: int = 6, m_ineq: int | None = None, sigma: float = 1.0,
                    upper_frac: float = 0.1, seed: int = 1234, masked_zero_divisors: bool = False) -> QnProblem:
    """Builds one synthetic KKT system. `masked_zero_divisors` plants z=0 divisors on masked-out lanes, the case the
    reference's own unit test plants (tests/LinAlg/vectorTests.hpp:1187-1191)."""
    if m_ineq is None:
        m_ineq = m // 2
    m_eq = m - m_ineq
    rj = np.random.default_rng(seed)
    J = rj.standard_normal((m, n)) / np.sqrt(n)
    if m > 0:
        J[0, :] = 1.0
    rv = np.random.default_rng(seed + 1)
    ixl = np.ones(n)
    ixu = (rv.random(n) < upper_frac).astype(np.float64)
    idl = np.ones(m_ineq)
    idu = (rv.random(m_ineq) < upper_frac).astype(np.float64)
    U = lambda k: rv.uniform(1e-3, 1.0, k)
    sxl, sxu, zl, zu = U(n), U(n), U(n), U(n)
    sdl, sdu, vl, vu = U(m_ineq), U(m_ineq), U(m_ineq), U(m_ineq)
    # the reference keeps duals/slacks of absent bounds at zero (matchesPattern asserts, hiopHessianLowRank.cpp:271-274)
    sxu *= ixu
    zu *= ixu
    sdu *= idu
    vu *= idu
    if masked_zero_divisors:
        pass  # sxu/sdu are already exactly 0 on masked-out lanes
    else:
        sxu[ixu == 0] = 1.0   # harmless non-zero divisors
        sdu[idu == 0] = 1.0
    St = rv.standard_normal((l, n))
    Yt = St * rv.uniform(0.5, 2.0, (l, n))
    L, D = secant_LD(St, Yt)
    rr = np.random.default_rng(seed + 2)
    rx, ryc, ryd = rr.standard_normal(n), rr.standard_normal(m_eq), rr.standard_normal(m_ineq)
    res = dict(rx=rx, rd=rr.standard_normal(m_ineq), ryc=ryc, ryd=ryd,
               rxl=rr.standard_normal(n) * ixl, rxu=rr.standard_normal(n) * ixu,
               rdl=rr.standard_normal(m_ineq) * idl, rdu=rr.standard_normal(m_ineq) * idu,
               rszl=rr.standard_normal(n) * ixl, rszu=rr.standard_normal(n) * ixu,
               rsvl=rr.standard_normal(m_ineq) * idl, rsvu=rr.standard_normal(m_ineq) * idu)
    return QnProblem(n=n, m_eq=m_eq, m_ineq=m_ineq, l=l, sigma=sigma, Jc=np.ascontiguousarray(J[:m_eq]),
                     Jd=np.ascontiguousarray(J[m_eq:]), ixl=ixl, ixu=ixu, idl=idl, idu=idu, sxl=sxl, sxu=sxu, zl=zl,
                     zu=zu, sdl=sdl, sdu=sdu, vl=vl, vu=vu, St=St, Yt=Yt, L=L, D=D, rx=rx, ryc=ryc, ryd=ryd, res=res)


def make_symmetric_indefinite(N: int, n_neg: int, seed: int = 7, cond: float = 1e3) -> np.ndarray:
    """Dense symmetric matrix with exactly n_neg negative eigenvalues (KKT-like inertia), full storage."""
    r = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(r.standard_normal((N, N)))
    ev = np.exp(r.uniform(0, np.log(cond), N))
    ev[:n_neg] *= -1.0
    M = (Q * ev) @ Q.T
    return 0.5 * (M + M.T)


def make_kkt_like(nx: int, m: int, seed: int = 11) -> np.ndarray:
    """[[H + D, J^T], [J, -Dd^{-1}]] with H SPD: inertia (nx, 0, m), the shape the MDS/dense-Newton paths factor
    (src/Optimization/hiopKKTLinSysMDS.cpp:172-305). Full symmetric storage."""
    r = np.random.default_rng(seed)
    A = r.standard_normal((nx, nx)) / np.sqrt(nx)
    H = A @ A.T + np.diag(r.uniform(1e-2, 1.0, nx))
    J = r.standard_normal((m, nx)) / np.sqrt(nx)
    K = np.zeros((nx + m, nx + m))
    K[:nx, :nx] = H
    K[nx:, :nx] = J
    K[:nx, nx:] = J.T
    K[nx:, nx:] = -np.diag(r.uniform(1e-3, 1.0, m))
    return K


@dataclass
class MdsProblem:
    """One synthetic mixed dense-sparse Newton KKT system (the shape hiopKKTLinSysCompressedMDSXYcYd assembles,
    src/Optimization/hiopKKTLinSysMDS.cpp:172-305): diagonal sparse Hessian block, dense H_d, sparse + dense Jacobian blocks."""
    nxs: int
    nxd: int
    neq: int
    nineq: int
    Hd: np.ndarray          # nxd x nxd symmetric
    Hs_diag: np.ndarray     # nxs
    Jcd: np.ndarray         # neq x nxd
    Jdd: np.ndarray         # nineq x nxd
    iRow_c: np.ndarray
    jCol_c: np.ndarray
    Jcs_vals: np.ndarray
    iRow_d: np.ndarray
    jCol_d: np.ndarray
    Jds_vals: np.ndarray
    ixl: np.ndarray
    ixu: np.ndarray
    idl: np.ndarray
    idu: np.ndarray
    sxl: np.ndarray
    sxu: np.ndarray
    zl: np.ndarray
    zu: np.ndarray
    sdl: np.ndarray
    sdu: np.ndarray
    vl: np.ndarray
    vu: np.ndarray
    delta_wx: np.ndarray
    delta_wd: np.ndarray
    delta_cc: np.ndarray
    delta_cd: np.ndarray
    rx: np.ndarray
    ryc: np.ndarray
    ryd: np.ndarray


def _sorted_triplets(r, m, n, nnz_per_row):
    rows, cols = [], []
    for i in range(m):
        k = min(n, nnz_per_row)
        cs = np.sort(r.choice(n, k, replace=False)) if n else np.zeros(0, dtype=int)
        rows += [i] * len(cs)
        cols += list(cs)
    return np.array(rows, dtype=np.int32), np.array(cols, dtype=np.int32)


def make_mds_problem(nxs: int, nxd: int, neq: int, nineq: int, nnz_per_row: int = 5, seed: int = 42, dwx: float = 0.0, dcc: float = 0.0) -> MdsProblem:
    r = np.random.default_rng(seed)
    n = nxs + nxd
    A = r.standard_normal((nxd, nxd)) / np.sqrt(max(nxd, 1))
    Hd = A @ A.T + np.diag(r.uniform(1e-2, 1.0, nxd))
    Hs = r.uniform(0.1, 2.0, nxs)
    Jcd = r.standard_normal((neq, nxd)) / np.sqrt(max(nxd, 1))
    Jdd = r.standard_normal((nineq, nxd)) / np.sqrt(max(nxd, 1))
    iRc, jCc = _sorted_triplets(r, neq, nxs, nnz_per_row)
    iRd, jCd = _sorted_triplets(r, nineq, nxs, nnz_per_row)
    ixl = np.ones(n)
    ixu = (r.random(n) < 0.2).astype(np.float64)
    idl = np.ones(nineq)
    idu = (r.random(nineq) < 0.2).astype(np.float64)
    U = lambda k: r.uniform(1e-3, 1.0, k)
    return MdsProblem(nxs=nxs, nxd=nxd, neq=neq, nineq=nineq, Hd=Hd, Hs_diag=Hs, Jcd=Jcd, Jdd=Jdd, iRow_c=iRc, jCol_c=jCc,
                      Jcs_vals=r.standard_normal(iRc.size), iRow_d=iRd, jCol_d=jCd, Jds_vals=r.standard_normal(iRd.size), ixl=ixl, ixu=ixu,
                      idl=idl, idu=idu, sxl=U(n), sxu=U(n) * ixu, zl=U(n), zu=U(n) * ixu, sdl=U(nineq), sdu=U(nineq) * idu, vl=U(nineq),
                      vu=U(nineq) * idu, delta_wx=np.full(n, dwx), delta_wd=np.full(nineq, dwx), delta_cc=np.full(neq, dcc),
                      delta_cd=np.full(nineq, dcc), rx=r.standard_normal(n), ryc=r.standard_normal(neq), ryd=r.standard_normal(nineq))


def make_secant_sequence(n, m_eq, m_ineq, steps=8, seed=31):
    """Iterate sequence for hiopHessianLowRank::update: x_k, grad_f_k (= diag(a) x_k, so s^T y > 0 on regular steps), multipliers,
    slowly varying Jacobians. Step 3 repeats the previous x (||s|| = 0 -> skipped), step 5 flips the gradient's sign against the
    step (s^T y < 0 -> skipped)."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(0.5, 2.0, n)
    Jc0, Jd0 = rng.standard_normal((m_eq, n)) / np.sqrt(n), rng.standard_normal((m_ineq, n)) / np.sqrt(n)
    seq = []
    x = rng.standard_normal(n)
    for k in range(steps):
        if k == 3:
            xk = seq[-1]["x"].copy()
        else:
            xk = x + 0.1 * rng.standard_normal(n)
        x = xk
        g = a * xk
        if k == 5:
            g = seq[-1]["grad_f"] - 3.0 * a * (xk - seq[-1]["x"])
        seq.append(dict(x=xk, grad_f=g, yc=rng.standard_normal(m_eq), yd=rng.standard_normal(m_ineq),
                        Jc=Jc0 + 1e-3 * k * rng.standard_normal((m_eq, n)) / np.sqrt(n),
                        Jd=Jd0 + 1e-3 * k * rng.standard_normal((m_ineq, n)) / np.sqrt(n)))
    return seq


def make_iterate(p: "QnProblem", seed: int = 61):
    """A full primal-dual iterate + NLP data around a QnProblem: the 12 blocks of hiopIterate, constraint values, gradient and
    bounds, as hiopResidual::update consumes them. Slack / dual blocks reuse the problem's (pattern-consistent) ones."""
    r = np.random.default_rng(seed)
    n, me, mi = p.n, p.m_eq, p.m_ineq
    itr = dict(x=r.standard_normal(n), d=r.standard_normal(mi), yc=r.standard_normal(me), yd=r.standard_normal(mi), sxl=p.sxl, sxu=p.sxu,
               sdl=p.sdl, sdu=p.sdu, zl=p.zl * p.ixl, zu=p.zu * p.ixu, vl=p.vl * p.idl, vu=p.vu * p.idu)
    data = dict(c=r.standard_normal(me), d=r.standard_normal(mi), grad=r.standard_normal(n), xl=np.where(p.ixl == 1.0, r.standard_normal(n), -1e20),
                xu=np.where(p.ixu == 1.0, 5.0 + r.standard_normal(n), 1e20), dl=np.where(p.idl == 1.0, r.standard_normal(mi), -1e20),
                du=np.where(p.idu == 1.0, 0.2 * r.standard_normal(mi), 1e20), crhs=r.standard_normal(me))
    return itr, data
