#!/usr/bin/env python
"""Dense symmetric factor+solve sweep (BASELINE configs[4]): hb_symdense in its three modes vs LAPACK DSYTRF/DPOTRF on the
host cores of the same box. Flops counted as N^3/3 like the reference does (FLOPS_DPOTRF, hiopLinSolverSymDenseMagma.cpp:155).
Usage: python tools/bench_symdense.py [N ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiop_b200 import synth  # noqa: E402
from hiop_b200.engine import Context, LinSolverSymDense  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192]
    ctx = Context(0)
    out = []
    for N in sizes:
        nx = (2 * N) // 3
        K = synth.make_kkt_like(nx, N - nx, seed=N)
        r = np.random.default_rng(0)
        A = r.standard_normal((N, N))
        S = A @ A.T + N * np.eye(N)
        rhs = r.standard_normal(N)
        row = {"N": N}
        for name, mode, M in (("bk", LinSolverSymDense.BUNCH_KAUFMAN, K), ("nopiv", LinSolverSymDense.NOPIV, K), ("chol", LinSolverSymDense.CHOLESKY, S)):
            s = LinSolverSymDense(ctx, N, mode)
            Md = ctx.to_device(np.triu(M))
            x = ctx.to_device(rhs)
            with ctx:
                ts, tsol = [], []
                for rep in range(3):
                    s.set_matrix(Md)
                    ctx.sync()
                    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    e0.record()
                    ret = s.matrixChanged()
                    e1.record()
                    x.copy_(torch.from_numpy(rhs))
                    s.solve(x)
                    e2.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                    tsol.append(e1.elapsed_time(e2))
            xs = x.cpu().numpy()
            res = np.abs(M @ xs - rhs).max() / np.abs(rhs).max()
            row[name] = {"factor_ms": min(ts), "solve_ms": min(tsol), "gflops": N ** 3 / 3 / min(ts) / 1e6, "ret": ret, "resid": res}
            s.close()
        from scipy.linalg import lapack
        t0 = time.perf_counter()
        ldu, piv, info = lapack.dsytrf(np.asfortranarray(np.tril(K)), lower=1, lwork=64 * N)
        row["lapack_dsytrf_ms"] = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        c, info = lapack.dpotrf(np.asfortranarray(S), lower=1)
        row["lapack_dpotrf_ms"] = (time.perf_counter() - t0) * 1e3
        row["host_cores"] = os.cpu_count()
        print(json.dumps(row))
        out.append(row)
    ctx.close()


if __name__ == "__main__":
    main()
