#!/usr/bin/env python
"""Dense symmetric factor+solve sweep (BASELINE configs[4]): hb_symdense in its three modes, next to cuSOLVER (through torch.linalg:
cholesky_ex = cusolverDnXpotrf, ldl_factor_ex = cusolverDnDsytrf -- the stand-in for the reference's MAGMA calls
src/LinAlg/hiopLinSolverSymDenseMagma.cpp:151,250,349,455) and LAPACK DSYTRF/DPOTRF on the host cores of the same box.
Flops counted as N^3/3 like the reference does (FLOPS_DPOTRF, hiopLinSolverSymDenseMagma.cpp:155).
Usage: python tools/bench_symdense.py [--no-lapack] [--modes bk,nopiv,chol] [N ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiop_b200 import synth  # noqa: E402
from hiop_b200.engine import Context, LinSolverSymDense  # noqa: E402


def _time(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    args = sys.argv[1:]
    no_lapack = "--no-lapack" in args
    modes = ("bk", "nopiv", "chol")
    if "--modes" in args:
        modes = tuple(args[args.index("--modes") + 1].split(","))
        del args[args.index("--modes"):args.index("--modes") + 2]
    sizes = [int(a) for a in args if not a.startswith("--")] or [512, 1024, 2048, 4096, 8192]
    ctx = Context(0)
    for N in sizes:
        nx = (2 * N) // 3
        r = np.random.default_rng(0)
        with ctx:
            # built on the device: the host versions take minutes at N = 24003
            g = torch.Generator(device="cuda").manual_seed(N)
            A = torch.randn(nx, nx, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(nx)
            H = A @ A.T + torch.diag(torch.rand(nx, dtype=torch.float64, device="cuda", generator=g) * 0.99 + 1e-2)
            J = torch.randn(N - nx, nx, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(nx)
            Kd = torch.zeros(N, N, dtype=torch.float64, device="cuda")
            Kd[:nx, :nx] = H
            Kd[nx:, :nx] = J
            Kd[:nx, nx:] = J.T
            Kd[nx:, nx:] = -torch.diag(torch.rand(N - nx, dtype=torch.float64, device="cuda", generator=g) * 0.999 + 1e-3)
            B = torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g)
            Sd = B @ B.T + N * torch.eye(N, dtype=torch.float64, device="cuda")
            del A, H, J, B
            rhs_d = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
            torch.cuda.synchronize()
        row = {"N": N}
        for name, mode, M in (("bk", LinSolverSymDense.BUNCH_KAUFMAN, Kd), ("nopiv", LinSolverSymDense.NOPIV, Kd), ("chol", LinSolverSymDense.CHOLESKY, Sd)):
            if name not in modes:
                continue
            s = LinSolverSymDense(ctx, N, mode)
            Md = torch.triu(M)
            x = rhs_d.clone()
            with ctx:
                ts, tsol = [], []
                for rep in range(3):
                    s.set_matrix(Md)
                    ctx.sync()
                    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    e0.record()
                    ret = s.matrixChanged()
                    e1.record()
                    x.copy_(rhs_d)
                    s.solve(x)
                    e2.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                    tsol.append(e1.elapsed_time(e2))
                res = float((M @ x - rhs_d).abs().max() / rhs_d.abs().max())
            row[name] = {"factor_ms": min(ts), "solve_ms": min(tsol), "tflops": N ** 3 / 3 / min(ts) / 1e9, "ret": ret, "resid": res}
            s.close()
            del Md
        # cuSOLVER through torch.linalg (device-resident, lower triangle), same matrices
        try:
            t_potrf = _time(lambda: torch.linalg.cholesky_ex(Sd))
            L = torch.linalg.cholesky_ex(Sd).L
            t_potrs = _time(lambda: torch.cholesky_solve(rhs_d[:, None], L))
            row["cusolver_potrf_ms"] = t_potrf
            row["cusolver_potrf_tflops"] = N ** 3 / 3 / t_potrf / 1e9
            row["cusolver_potrs_ms"] = t_potrs
            del L
        except Exception as e:  # noqa: BLE001
            row["cusolver_potrf_error"] = str(e)[:100]
        try:
            t_sytrf = _time(lambda: torch.linalg.ldl_factor_ex(Kd), reps=2)
            LD, piv, _ = torch.linalg.ldl_factor_ex(Kd)
            t_sytrs = _time(lambda: torch.linalg.ldl_solve(LD, piv, rhs_d[:, None]), reps=2)
            row["cusolver_sytrf_ms"] = t_sytrf
            row["cusolver_sytrf_tflops"] = N ** 3 / 3 / t_sytrf / 1e9
            row["cusolver_sytrs_ms"] = t_sytrs
            del LD, piv
        except Exception as e:  # noqa: BLE001
            row["cusolver_sytrf_error"] = str(e)[:100]
        if not no_lapack and N <= 8192:
            from scipy.linalg import lapack
            K = Kd.cpu().numpy()
            S = Sd.cpu().numpy()
            t0 = time.perf_counter()
            lapack.dsytrf(np.asfortranarray(np.tril(K)), lower=1, lwork=64 * N)
            row["lapack_dsytrf_ms"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            lapack.dpotrf(np.asfortranarray(S), lower=1)
            row["lapack_dpotrf_ms"] = (time.perf_counter() - t0) * 1e3
            row["host_cores"] = os.cpu_count()
        print(json.dumps(row), flush=True)
        del Kd, Sd
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
