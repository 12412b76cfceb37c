#!/usr/bin/env python
"""One factor + solve of the B1 solver for profiling (ncu launch lists / --set full captures).
Usage: python tools/prof_symdense.py MODE N [reps]   (MODE = bk | nopiv | chol)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiop_b200.engine import Context, LinSolverSymDense  # noqa: E402


def main():
    mode_name, N = sys.argv[1], int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    mode = {"bk": LinSolverSymDense.BUNCH_KAUFMAN, "nopiv": LinSolverSymDense.NOPIV, "chol": LinSolverSymDense.CHOLESKY}[mode_name]
    ctx = Context(0)
    g = torch.Generator(device="cuda").manual_seed(N)
    if mode_name == "chol":
        B = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g)
        M = B @ B.T + N * torch.eye(N, dtype=torch.float64, device="cuda")
    else:
        nx = (2 * N) // 3
        A = torch.randn(nx, nx, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(nx)
        M = torch.zeros(N, N, dtype=torch.float64, device="cuda")
        M[:nx, :nx] = A @ A.T + torch.diag(torch.rand(nx, dtype=torch.float64, device="cuda", generator=g) * 0.99 + 1e-2)
        J = torch.randn(N - nx, nx, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(nx)
        M[nx:, :nx] = J
        M[:nx, nx:] = J.T
        M[nx:, nx:] = -torch.diag(torch.rand(N - nx, dtype=torch.float64, device="cuda", generator=g) * 0.999 + 1e-3)
    rhs = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    torch.cuda.synchronize()
    s = LinSolverSymDense(ctx, N, mode)
    Mu = torch.triu(M)
    for _ in range(reps):
        with ctx:
            s.set_matrix(Mu)
            ret = s.matrixChanged()
            x = rhs.clone()
            s.solve(x)
            ctx.sync()
    print("ret", ret, "resid", float((M @ x - rhs).abs().max() / rhs.abs().max()))
    s.close()
    ctx.close()


if __name__ == "__main__":
    main()
