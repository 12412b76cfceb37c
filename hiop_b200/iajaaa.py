"""HiOp's .iajaaa KKT dump format (src/LinAlg/csr_iajaaa.md, writer src/Utils/hiopCSR_IO.hpp:44-155, Matlab reader
src/LinAlg/load_kkt_mat.m): host-side helpers around the C-ABI writer + a reader, so KKT systems can be exchanged with upstream
HiOp builds (`write_kkt yes`) and replayed through the engine."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def write_system(filename: str, M_upper: np.ndarray, nx: int, meq: int, mineq: int, pairs=()):
    """Matrix (N x N row-major, upper triangle used) followed by (rhs, solution) vectors, exactly like the reference writer."""
    L = _lib.lib()
    M = np.ascontiguousarray(M_upper, dtype=np.float64)
    N = M.shape[0]
    dp = ctypes.POINTER(ctypes.c_double)
    rc = L.hb_iajaaa_write_matrix_host(filename.encode(), N, M.ctypes.data_as(dp), int(nx), int(meq), int(mineq))
    if rc != 0:
        raise RuntimeError(L.hb_last_error().decode())
    for rhs, sol in pairs:
        for v in (rhs, sol):
            v = np.ascontiguousarray(v, dtype=np.float64)
            rc = L.hb_iajaaa_append_vector_host(filename.encode(), N, v.ctypes.data_as(dp))
            if rc != 0:
                raise RuntimeError(L.hb_last_error().decode())


def read_system(filename: str):
    """-> dict(N, nx, meq, mineq, M (dense, upper triangle), pairs=[(rhs, sol), ...])."""
    with open(filename) as f:
        tok = f.read().split()
    N, nx, meq, mineq, nnz = (int(t) for t in tok[:5])
    o = 5
    ia = np.array(tok[o:o + N + 1], dtype=np.int64); o += N + 1
    ja = np.array(tok[o:o + nnz], dtype=np.int64); o += nnz
    aa = np.array(tok[o:o + nnz], dtype=np.float64); o += nnz
    M = np.zeros((N, N))
    for i in range(N):
        for p in range(ia[i] - 1, ia[i + 1] - 1):
            M[i, ja[p] - 1] = aa[p]
    vecs = []
    while o + N <= len(tok):
        vecs.append(np.array(tok[o:o + N], dtype=np.float64)); o += N
    pairs = [(vecs[i], vecs[i + 1]) for i in range(0, len(vecs) - 1, 2)]
    return dict(N=N, nx=nx, meq=meq, mineq=mineq, M=M, pairs=pairs)
