import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hiop_b200 import synth
from hiop_b200.engine import Context, LinSolverSymDense
ctx = Context(0)
for nx, m in ((1500, 1001), (1500, 1000), (1500, 1002), (1400, 900), (2200, 900), (1000, 1001)):
    K = synth.make_kkt_like(nx, m, seed=nx + m)
    N = nx + m
    s = LinSolverSymDense(ctx, N, LinSolverSymDense.BUNCH_KAUFMAN)
    s.set_matrix(ctx.to_device(np.triu(K)))
    ret = s.matrixChanged()
    rhs = np.random.default_rng(0).standard_normal(N)
    x = ctx.to_device(rhs.copy())
    s.solve(x); ctx.sync()
    xs = x.cpu().numpy()
    print(nx, m, "ret", ret, "resid", np.abs(K @ xs - rhs).max(), "nan", np.isnan(xs).sum())
    s.close()
