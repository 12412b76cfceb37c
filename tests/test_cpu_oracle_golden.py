"""CPU suite: the oracle restatement against the committed golden fixtures (produced by the UNMODIFIED reference,
tests/golden/make_golden.py), plus C-ABI load/export checks. No GPU needed."""
import glob
import os

import numpy as np
import pytest

from hiop_b200 import _lib
from oracle import kkt_oracle as ko

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLD, name)))


def _state(g):
    Dx, DhInv, Dd, Dd_inv = ko.kkt_update(g["zl"], g["sxl"], g["zu"], g["sxu"], g["ixl"], g["ixu"], g["vl"], g["sdl"], g["vu"],
                                          g["sdu"], g["idl"], g["idu"], float(g["sigma"]))
    st = ko.QnState(g["Jc"], g["Jd"], DhInv, Dd_inv, g["St"], g["Yt"], g["L"], g["D"], float(g["sigma"]))
    return Dx, st


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "qn_*.npz"))))
def test_oracle_qn_against_golden(name):
    g = _load(name)
    Dx, st = _state(g)
    np.testing.assert_array_equal(Dx, g["ref_Dx"])            # elementwise: bit-exact
    np.testing.assert_array_equal(st.DhInv, g["ref_DhInv"])
    np.testing.assert_array_equal(st.Dd_inv, g["ref_Dd_inv"])
    N, _, _, _ = ko.condense(st)
    assert np.abs(N - g["ref_N"]).max() <= 1e-12 * np.abs(g["ref_N"]).max()
    hs = ko.hess_solve(st, g["rx"])
    assert np.abs(hs - g["ref_hess_solve"]).max() <= 1e-11 * np.abs(g["ref_hess_solve"]).max()
    dx, dyc, dyd, _ = ko.solve_compressed(st, g["rx"], g["ryc"], g["ryd"])
    for a, b in ((dx, g["ref_dx"]), (dyc, g["ref_dyc"]), (dyd, g["ref_dyd"])):
        if b.size:
            assert np.abs(a - b).max() <= 1e-8 * max(1.0, np.abs(b).max())
    res = {k: g["res_" + k] for k in ko.RES_NAMES}
    it = {k: g[k] for k in ("sxl", "sxu", "zl", "zu", "sdl", "sdu", "vl", "vu")}
    pat = {k: g[k] for k in ("ixl", "ixu", "idl", "idu")}
    d = ko.compute_directions(st, it, pat, res)
    for k in ko.DIR_NAMES:
        b = g["ref_dir_" + k]
        assert np.abs(d[k] - b).max() <= 1e-8 * max(1.0, np.abs(b).max()) if b.size else True, k
    Bx = ko.hess_times_vec(g["St"], g["Yt"], float(g["sigma"]), Dx, 0.0, np.zeros_like(Dx), 1.0, g["tv_x"], True)
    assert np.abs(Bx - g["ref_Bx"]).max() <= 1e-10 * np.abs(g["ref_Bx"]).max()
    # outer refinement: full 12-block operator and compute_directions_w_IR
    y = ko.kkt_full_times_vec(st, it, pat, {k: g["kx_in_" + k] for k in ko.DIR_NAMES}, Dx)
    for k in ko.RES_NAMES:
        b = g["ref_kx_out_" + k]
        tol = 1e-12 if k in ("rx", "ryc", "ryd") else 0.0
        assert np.abs(y[k] - b).max(initial=0.0) <= tol * max(1.0, np.abs(b).max(initial=0.0)), k
    d, info = ko.compute_directions_w_ir(st, it, pat, res, float(g["ir_mu"]), int(g["ir_maxit"]), Dx=Dx)
    assert info[0] == g["ref_ir_info"][0] and info[1] == g["ref_ir_info"][1]
    for k in ko.DIR_NAMES:
        b = g["ref_ir_dir_" + k]
        assert np.abs(d[k] - b).max(initial=0.0) <= 1e-8 * max(1.0, np.abs(b).max(initial=0.0)), k


def test_oracle_symdense_against_golden():
    g = _load("symdense.npz")
    for i in range(int(g["count"])):
        ret, f = ko.symdense_matrix_changed(g[f"K{i}"])
        assert ret == int(g[f"ret{i}"]), i
        if ret >= 0:
            sol = f.solve(g[f"rhs{i}"])
            assert np.abs(sol - g[f"sol{i}"]).max() <= 1e-9 * max(1.0, np.abs(g[f"sol{i}"]).max()), i


def test_oracle_vector_ops_against_golden():
    g = _load("vector_ops.npz")
    y, x, z, sel, ixu = g["y"], g["x"], g["z"], g["sel"], g["ixu"]
    z0 = z * sel
    for alpha in (1.0, -1.0, 0.37):
        np.testing.assert_array_equal(ko.axdzpy_w_pattern(y.copy(), alpha, x, z0, sel), g[f"axdzpy_w_pattern_{alpha}"])
        np.testing.assert_array_equal(ko.axzpy(y.copy(), alpha, x, z), g[f"axzpy_{alpha}"])
    np.testing.assert_array_equal(ko.component_div_w_select(y.copy(), z0, sel), g["component_div_w_sel"])
    np.testing.assert_array_equal(ko.add_log_barrier_grad(y.copy(), 0.1, z0, sel), g["add_logbar_grad"])
    np.testing.assert_array_equal(ko.add_linear_damping_term(y.copy(), sel, ixu, 0.9, 1e-6), g["add_lin_damping"])
    assert ko.log_barrier(z, sel) == float(g["logbarrier"])
    assert ko.linear_damping_term(z, sel, ixu, 0.1, 1e-5) == float(g["lin_damping_term"])
    assert ko.fraction_to_the_bdry(z, x, 0.995) == float(g["frac_to_bdry"])
    assert ko.fraction_to_the_bdry(z, x, 0.995, sel) == float(g["frac_to_bdry_w_sel"])


def test_iajaaa_writer_against_reference_golden(tmp_path):
    """The .iajaaa writer of the C-ABI (host entry points, no GPU needed) reproduces the reference's file byte for byte."""
    from hiop_b200 import iajaaa
    g = _load("iajaaa_case.npz")
    out = str(tmp_path / "k.iajaaa")
    iajaaa.write_system(out, g["K"], int(g["nx"]), int(g["meq"]), int(g["mineq"]), [(g["rhs"], g["sol"])])
    gold = os.path.join(os.path.dirname(__file__), "golden", "kkt_linsys_7.iajaaa")
    assert open(out, "rb").read() == open(gold, "rb").read()
    back = iajaaa.read_system(gold)
    assert back["N"] == 32 and len(back["pairs"]) == 1


def test_cabi_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    syms = _lib.declared_symbols()
    assert len(syms) >= 60
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert b"sm_100a" in L.hb_version()


def test_ctypes_prototypes_match_the_header():
    """Every function of include/hiopb200.h has a ctypes prototype with the same number of parameters and a compatible scalar /
    pointer kind per position (a wrong count or a double passed where the C side expects a pointer would only show up on a GPU)."""
    import ctypes
    import re
    L = _lib.lib()
    txt = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)
    protos = re.findall(r"\b(?:int|long long|double|const char\*|const double\*|void)\s*\*?\s*(hb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S)
    assert len(protos) >= 90
    bad = []
    for name, params in protos:
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [q.strip() for q in params.split(",")]
        fn = getattr(L, name)
        if fn.argtypes is None:
            bad.append((name, "no ctypes prototype"))
            continue
        if len(fn.argtypes) != len(plist):
            bad.append((name, f"header has {len(plist)} parameters, ctypes {len(fn.argtypes)}"))
            continue
        for i, (cdecl, at) in enumerate(zip(plist, fn.argtypes)):
            is_ptr_c = "*" in cdecl
            is_ptr_py = at in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(at, "contents") or issubclass(at, ctypes._Pointer)
            if is_ptr_c != is_ptr_py:
                bad.append((name, f"parameter {i} '{cdecl}' vs {at.__name__}"))
            elif not is_ptr_c:
                want = ctypes.c_double if cdecl.split()[0] == "double" else None
                if want is not None and at is not ctypes.c_double:
                    bad.append((name, f"parameter {i} '{cdecl}' vs {at.__name__}"))
                if want is None and at is ctypes.c_double:
                    bad.append((name, f"parameter {i} '{cdecl}' vs {at.__name__}"))
    assert not bad, bad


def test_engine_fails_loudly_without_gpu():
    import ctypes
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    h = ctypes.c_void_p()
    rc = L.hb_ctx_create(0, ctypes.byref(h))
    assert rc == -2 and b"no CPU fallback" in L.hb_last_error()
    from hiop_b200.engine import Context
    with pytest.raises(_lib.EngineError):
        Context(0)
