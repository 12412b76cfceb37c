"""Drop-in parity on the reference's own bundled drivers (BASELINE configs[0] and the iterate-sequence gate).

oracle/_ref/{ex1,ex2,mds1}_b200.exe are the UNMODIFIED reference drivers linked against the reference library with the
two factory hooks of INTEGRATION.md applied; HIOP_B200=1 routes hiopKKTLinSysLowRank::{update,solveCompressed}
(quasi-Newton drivers) / hiopLinSolverSymDense::{matrixChanged,solve} (MDS driver) to libhiopb200.so, HIOP_B200 unset
runs the reference's CPU LAPACK classes through the same binary. The iteration tables must agree column by column
within 1e-5 -- the tolerance the reference itself uses between its CPU and RAJA builds
(tests/testMDS1CompareIterations.awk:13) -- with equal iteration counts, and the -selfcheck objectives must pass."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
ROW = re.compile(r"^\s*(\d+)\s+([-+]?\d\.\d+e[-+]\d+)\s+(\d\.\d+e[-+]\d+)\s+(\d\.\d+e[-+]\d+)\s+(-?\d+\.\d+)\s+(\d\.\d+e[-+]\d+)\s+(\d\.\d+e[-+]\d+)")


def _run(exe, args, b200):
    path = os.path.join(REF, exe)
    if not os.path.exists(path):
        pytest.skip(f"{exe} not built (oracle/_ref travels from the build container)")
    env = dict(os.environ)
    env.pop("HIOP_B200", None)
    if b200:
        env["HIOP_B200"] = "1"
    p = subprocess.run([path] + args, capture_output=True, text=True, env=env, timeout=600)
    table = []
    warned = False
    for line in p.stdout.splitlines():
        m = ROW.match(line)
        if m:
            # last entry: 1.0 when a "solveWithRefin reduced residual to ONLY ..." warning preceded this iterate,
            # i.e. the step that produced it came from a solve that did not reach the 1e-8 refinement tolerance
            table.append([float(x) for x in m.groups()] + [1.0 if warned else 0.0])
            warned = False
        elif "reduced residual to ONLY" in line:
            warned = True
    return p.returncode, p.stdout, table


def _compare(exe, args, inf_du_rel=0.0):
    """Column-by-column comparison with the reference's own rule: |a-b| <= 1e-5 ABSOLUTE on every printed column
    (tests/testMDS1CompareIterations.awk:13,26), equal iteration counts. The objective is additionally held to 1e-7
    relative. With `inf_du_rel` > 0 the inf_du column of the iterates that the REFERENCE itself flags as coming from an
    unconverged solveWithRefin must only be no worse than the reference's (see test_ex2_*)."""
    rc_r, out_r, tab_r = _run(exe, args, False)
    rc_b, out_b, tab_b = _run(exe, args, True)
    assert rc_r == 0, out_r[-2000:]
    assert rc_b == 0, out_b[-2000:]
    assert len(tab_r) > 3
    assert len(tab_b) == len(tab_r), (len(tab_b), len(tab_r))
    worst = 0.0
    for a, b in zip(tab_b, tab_r):
        assert a[0] == b[0]
        assert abs(a[1] - b[1]) <= 1e-7 * max(1.0, abs(b[1])), (a, b)
        for j in (2, 3, 4, 5, 6):   # inf_pr, inf_du, lg(mu), alpha_du, alpha_pr
            d = abs(a[j] - b[j])
            if j == 3 and inf_du_rel > 0 and b[7] == 1.0 and a[j] <= (1.0 + inf_du_rel) * b[j]:
                continue
            worst = max(worst, d)
    return worst, len(tab_r), out_b


@pytest.fixture(params=["device", "host"])
def ir_mode(request):
    """Where the outer BiCGStab refinement runs: on the device (hb_lowrank_compute_directions_w_ir, the default) or in the
    reference's host code calling the engine's solveCompressed for every preconditioner apply."""
    os.environ["HIOP_B200_IR"] = request.param
    yield request.param
    os.environ.pop("HIOP_B200_IR", None)


@pytest.mark.parametrize("args", [["500", "-selfcheck"], ["5000", "-selfcheck"], ["5000", "-unconstrained", "-selfcheck"]])
def test_ex2_iterate_sequence(args, ir_mode):
    # Constrained Ex2 drives the condensed matrix N to the edge of FP64: the REFERENCE's own solveWithRefin prints
    # "reduced residual to ONLY 7.6e-06 after 3 iterative refinements" before 22 of its 35 iterates at n=5000. The dual
    # step of such an iterate (hence the printed inf_du) carries a component from N's near-null space that depends on
    # the summation order of J Dx^-1 J^T, so it is not reproducible even between two CPU BLAS builds. Measured on B200
    # (tools/dropin_diff.sh): objective, alpha_pr, alpha_du, lg(mu) identical at every iteration, inf_pr within 7e-6
    # absolute, inf_du identical except at flagged iterates 3 (1.785e+02 vs the reference's 8.802e+02) and 7 (2.273e+01
    # vs 2.304e+01), where the engine's solve converged and the reference's did not. Rule: on flagged iterates inf_du
    # must be no worse than the reference's (+2%); everything else is held to the reference's 1e-5 absolute rule.
    rel = 0.0 if "-unconstrained" in args else 2e-2
    worst, nit, out = _compare("ex2_b200.exe", args, inf_du_rel=rel)
    assert "selfcheck success" in out
    assert worst <= 1e-5, worst


@pytest.mark.parametrize("args", [["500", "1.0", "-selfcheck"], ["1000", "1.0"], ["50000", "1.0", "-selfcheck"]])
def test_ex1_iterate_sequence(args, ir_mode):
    worst, nit, out = _compare("ex1_b200.exe", args)
    assert worst <= 1e-5, worst


@pytest.mark.parametrize("linsol", ["bk", "nopiv"])
def test_mds1_iterate_sequence(linsol):
    os.environ["HIOP_B200_LINSOL"] = linsol
    try:
        worst, nit, out = _compare("mds1_b200.exe", ["400", "100", "0", "-selfcheck"])
    finally:
        os.environ.pop("HIOP_B200_LINSOL", None)
    assert "selfcheck passed" in out
    assert worst <= 1e-5, worst


# ---------------------------------------------------------------------------------------------------------------------------------
# parametrised-m dense-constraints problem (oracle/drivers/NlpDenseConsExM.cpp: this repo's generalisation of NlpDenseConsEx1, SURVEY
# section 0 item 4): the bundled drivers have m <= 4, so only here does the default int8-slice condensation run inside the interior-point
# loop (AUTO switches to it for global n >= 32768 and m + 2l >= 64). Reference vs HB_CONDENSE=dmma vs HB_CONDENSE=oz8 under the 1e-5 rule.
# ---------------------------------------------------------------------------------------------------------------------------------
ROW_LS = re.compile(ROW.pattern + r"\s+(\d+)\(")


def _run_env(exe, args, env_extra, cwd=None):
    path = os.path.join(REF, exe)
    if not os.path.exists(path):
        pytest.skip(f"{exe} not built (oracle/_ref travels from the build container)")
    env = dict(os.environ)
    for k in ("HIOP_B200", "HB_CONDENSE", "HIOP_B200_STATS", "HIOP_B200_SECANT"):
        env.pop(k, None)
    env.update(env_extra)
    p = subprocess.run([path] + args, capture_output=True, text=True, env=env, timeout=900, cwd=cwd)
    table = []
    for line in p.stdout.splitlines():
        m = ROW_LS.match(line)
        if m:
            table.append([float(x) for x in m.groups()])        # ..., alpha_pr, number of line-search trials
        elif ROW.match(line):
            table.append([float(x) for x in ROW.match(line).groups()] + [0.0])
    return p.returncode, p.stdout, p.stderr, table


def _tables_agree(tab_b, tab_r, until_linesearch_differs=False):
    """1e-5 absolute on inf_pr, inf_du, lg(mu), alpha_du, alpha_pr and 1e-7 relative on the objective, row by row. With
    `until_linesearch_differs` the comparison stops at the first iterate whose number of line-search trials differs: a filter decision on
    the boundary flips with the summation order of the dot products (the reference's threaded OpenBLAS is itself not reproducible run to
    run there), and every later row differs legitimately. Returns (worst difference, rows compared)."""
    assert len(tab_r) > 3
    worst, rows = 0.0, 0
    for a, b in zip(tab_b, tab_r):
        if until_linesearch_differs and a[7] != b[7]:
            break
        assert a[0] == b[0]
        assert abs(a[1] - b[1]) <= 1e-7 * max(1.0, abs(b[1])), (a, b)
        worst = max(worst, max(abs(a[j] - b[j]) for j in (2, 3, 4, 5, 6)))
        rows += 1
    if not until_linesearch_differs:
        assert len(tab_b) == len(tab_r), (len(tab_b), len(tab_r))
    return worst, rows


@pytest.mark.parametrize("n,m", [(1000, 50), (33000, 64), (33000, 128)])
def test_exM_iterate_sequence_both_condensation_kernels(n, m):
    rc_r, out_r, _, tab_r = _run_env("exM_b200.exe", [str(n), str(m)], {})
    assert rc_r == 0, out_r[-1500:]
    obj_r = float(re.search(r"objective=([-+0-9.e]+)", out_r).group(1))
    tabs = {}
    for mode in ("dmma", "oz8", None):      # None = AUTO (int8 slices at the two large sizes, FP64 DMMA at n = 1000)
        env = {"HIOP_B200": "1"}
        if mode:
            env["HB_CONDENSE"] = mode
        rc_b, out_b, err_b, tab_b = _run_env("exM_b200.exe", [str(n), str(m)], env)
        assert rc_b == 0, (mode, out_b[-1500:], err_b[-500:])
        tabs[mode] = tab_b
        # against the reference: the 1e-5 rule on every iterate up to the first flipped line-search decision (late, mu ~ 1e-7), which must
        # not come early; same optimum; iteration counts within 2
        worst, rows = _tables_agree(tab_b, tab_r, until_linesearch_differs=True)
        assert worst <= 1e-5, (mode, worst)
        assert rows >= min(25, len(tab_r)), (mode, rows, len(tab_r))
        assert abs(len(tab_b) - len(tab_r)) <= 2, (mode, len(tab_b), len(tab_r))
        obj_b = float(re.search(r"objective=([-+0-9.e]+)", out_b).group(1))
        assert abs(obj_b - obj_r) <= 1e-8 * abs(obj_r), (mode, obj_b, obj_r)
    # the int8-slice condensation inside the interior-point loop reproduces the exact FP64 path of the engine (same rule: the host part
    # of the loop -- the reference's own secant update, residuals and line search on threaded BLAS -- is not bitwise reproducible)
    for mode in ("oz8", None):
        worst, rows = _tables_agree(tabs[mode], tabs["dmma"], until_linesearch_differs=True)
        assert worst <= 1e-5, (mode, worst)
        assert rows >= min(25, len(tabs["dmma"])), (mode, rows)


@pytest.mark.parametrize("exe,args", [("exM_b200.exe", ["33000", "64"]), ("exM_b200.exe", ["1000", "50"]), ("ex2_b200.exe", ["5000", "-unconstrained", "-selfcheck"]),
                                      ("ex1_b200.exe", ["50000", "1.0", "-selfcheck"])])
def test_secant_memory_on_the_device(exe, args):
    """a11 inside the drop-in (HIOP_B200_SECANT=device): hiopHessianLowRankB200::update only notes the iterate, the KKT adapter hands it to
    hb_lowrank_secant_update, S_t / Y_t / x_prev / grad_prev / J_prev never leave HBM. Same iterate table as the reference (1e-5 rule up to
    the first flipped line-search decision), same optimum, and every accepted iterate went through the device-side update."""
    rc_r, out_r, _, tab_r = _run_env(exe, args, {})
    assert rc_r == 0, out_r[-1500:]
    rc_b, out_b, err_b, tab_b = _run_env(exe, args, {"HIOP_B200": "1", "HIOP_B200_SECANT": "device", "HIOP_B200_STATS": "1"})
    assert rc_b == 0, (out_b[-1500:], err_b[-800:])
    m = re.search(r"updates (\d+), .*secant updates on the device (\d+)", err_b)
    assert m and int(m.group(2)) == int(m.group(1)) > 3, err_b[-800:]
    worst, rows = _tables_agree(tab_b, tab_r, until_linesearch_differs=True)
    assert worst <= 1e-5, worst
    assert rows >= min(25, len(tab_r)), (rows, len(tab_r))
    assert abs(len(tab_b) - len(tab_r)) <= 2, (len(tab_b), len(tab_r))
    if "-selfcheck" in args:                      # the drivers return non-zero when their selfcheck fails (rc checked above)
        assert "selfcheck" in out_b
    else:
        obj = [float(re.search(r"objective=([-+0-9.e]+)", o).group(1)) for o in (out_b, out_r)]
        assert abs(obj[0] - obj[1]) <= 1e-8 * abs(obj[1]), obj


def test_exM_jacobian_is_uploaded_once():
    """zero Jacobian bytes after the first upload: (a) the problem declares itself linear -> HiOp evaluates the Jacobian once and the adapter
    sees an unchanged evaluation counter; (b) it does not -> the (small) Jacobian is fingerprinted and found unchanged."""
    for extra in (["-linear"], []):
        rc, out, err, tab = _run_env("exM_b200.exe", ["4000", "60"] + extra, {"HIOP_B200": "1", "HIOP_B200_STATS": "1"})
        assert rc == 0, out[-1500:]
        m = re.search(r"Jacobian uploads (\d+) \(first (\d+) bytes, after the first (\d+) bytes\), KKT update\+condense ([\d.]+) ms/it", err)
        assert m, err[-800:]
        assert int(m.group(1)) == 1 and int(m.group(2)) == 8 * 60 * 4000 and int(m.group(3)) == 0, m.groups()
        assert len(tab) > 5


def test_speculative_mode_follows_the_kkt_safe_mode_flag(tmp_path):
    """linsol_mode=speculative (set through a hiop.options file in the working directory): HiOp starts with safe mode OFF, so the adapter
    factorizes with LDL^T without pivoting, and switches to Bunch-Kaufman whenever HiOp turns safe mode on -- the adapter reads the KKT
    object's flag at every matrixChanged() (the reference's non-MAGMA build never re-creates the solver on a flip). Same iterate table as
    the reference's LAPACK classes under the same option, selfcheck passes."""
    path = os.path.join(REF, "mds1_b200.exe")
    if not os.path.exists(path):
        pytest.skip("mds1_b200.exe not built")
    (tmp_path / "hiop.options").write_text("linsol_mode speculative\n")
    tabs = []
    for b200 in (False, True):
        env = dict(os.environ)
        for k in ("HIOP_B200", "HIOP_B200_LINSOL"):
            env.pop(k, None)
        if b200:
            env["HIOP_B200"] = "1"
        p = subprocess.run([path, "400", "100", "0", "-selfcheck"], capture_output=True, text=True, env=env, timeout=600, cwd=str(tmp_path))
        assert p.returncode == 0 and "selfcheck passed" in p.stdout, p.stdout[-1500:]
        tabs.append([[float(x) for x in m.groups()] for m in (ROW.match(line) for line in p.stdout.splitlines()) if m])
    assert _tables_agree([r + [0.0] for r in tabs[1]], [r + [0.0] for r in tabs[0]])[0] <= 1e-5
