#!/usr/bin/env python
"""Phase cycle counters of the dense-factor kernels that sit on the critical path (diagnostics):
  * k_diag128 (128 x 128 diagonal block of the look-ahead Cholesky / LDL^T): load, 16x16 factor+inverse, rows below, rank-16 update, store, inversion
  * k_bk_panel (cluster Bunch-Kaufman panel): load, column max, barrier, fail path, interchange, pivot, write-back
Usage: python tools/prof_diag.py [N]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiop_b200 import _lib  # noqa: E402
from hiop_b200.engine import Context, LinSolverSymDense  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    ctx = Context(0)
    L = _lib.lib()
    L.hb_debug_diag128_profile.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    L.hb_debug_bk_profile.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
    g = torch.Generator(device="cuda").manual_seed(N)
    B = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g)
    S = torch.triu(B @ B.T + N * torch.eye(N, dtype=torch.float64, device="cuda"))
    for ldl in (0, 1):
        s = LinSolverSymDense(ctx, N, LinSolverSymDense.CHOLESKY)
        with ctx:
            s.set_matrix(S)
            ctx.sync()
        prof = (ctypes.c_longlong * 8)()
        assert L.hb_debug_diag128_profile(s.h, ldl, prof) == 0
        names = ["load", "16x16 factor+inverse (x8)", "rows below (x8)", "rank-16 update (x8)", "store L", "128x128 inversion"]
        print(f"k_diag128<ldl={ldl}> cycles:", {n: int(prof[i]) for i, n in enumerate(names)}, "total", sum(prof[:6]))
        s.close()
    nx = (2 * N) // 3
    A = torch.randn(nx, nx, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(nx)
    K = torch.zeros(N, N, dtype=torch.float64, device="cuda")
    K[:nx, :nx] = A @ A.T + torch.diag(torch.rand(nx, dtype=torch.float64, device="cuda", generator=g) * 0.99 + 1e-2)
    J = torch.randn(N - nx, nx, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(nx)
    K[nx:, :nx] = J
    K[:nx, nx:] = J.T
    K[nx:, nx:] = -torch.diag(torch.rand(N - nx, dtype=torch.float64, device="cuda", generator=g) * 0.999 + 1e-3)
    s = LinSolverSymDense(ctx, N, LinSolverSymDense.BUNCH_KAUFMAN)
    with ctx:
        s.set_matrix(torch.triu(K))
        ctx.sync()
        assert L.hb_debug_bk_profile(ctx.h, 1, None) == 0
        ret = s.matrixChanged()
        prof = (ctypes.c_longlong * 8)()
        assert L.hb_debug_bk_profile(ctx.h, 0, prof) == 0
    names = ["load", "column max + post", "cluster barrier", "decide / fail path", "interchange", "pivot + update", "write-back"]
    print(f"k_bk_panel N={N} ret={ret} cycles (sum over panels, CTA 0):", {n: int(prof[i]) for i, n in enumerate(names)}, "total", sum(prof[:7]),
          "per column", sum(prof[:7]) // N)
    s.close()
    ctx.close()


if __name__ == "__main__":
    main()
