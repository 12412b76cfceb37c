"""CPU model of the int8-slice (Ozaki-scheme) condensation of hiop_b200/csrc/hb_ozaki.cu -- TEST INFRASTRUCTURE ONLY.

Restates, in numpy integer arithmetic, what the device path does so that its exactness claims can be checked without a GPU:
row exponents from the exact row maximum (k_oz_rowmax / k_oz_exponents), digits from two magic-number roundings and balanced base-128
extraction (k_oz_slice), exact int32 slice products with the truncation rule s + t <= S-1 and K chunks of 2^19 / S columns (k_oz_gemm),
FP64 recombination with the weights 2^-(12+7(s+t)) and the row scales (k_oz_fixup). It is not used by the product."""
from __future__ import annotations

import numpy as np

MAGIC = 6755399441055744.0  # 1.5 * 2^52


def row_exponents(B: np.ndarray) -> np.ndarray:
    """e_i with max_k |b_ik| = f * 2^e_i, f in [0.5, 1) (frexp); 0 for an all-zero row."""
    mx = np.abs(B).max(axis=1)
    e = np.zeros(B.shape[0], dtype=np.int64)
    nz = mx > 0
    e[nz] = np.frexp(mx[nz])[1]
    return e


def _digits(v: np.ndarray, nd: int) -> list[np.ndarray]:
    """Balanced base-128 digits of the integers v, least significant extracted first, leading digit = what is left."""
    v = v.astype(np.int64)
    d = [None] * nd
    for j in range(nd - 1, 0, -1):
        d[j] = ((v + 64) & 127) - 64
        v = (v - d[j]) >> 7
    d[0] = v
    return d


def slices(B: np.ndarray, e: np.ndarray, S: int) -> np.ndarray:
    """Q[p] (int64 holding int8 values): sum_p Q[p] 2^-(6+7p) = B / 2^e rounded to the last slice's grid."""
    assert 5 <= S <= 8
    nlo = S - 4
    xs = np.ldexp(B, (27 - e)[:, None])                      # |xs| < 2^27
    t = xs + MAGIC
    hi = (t - MAGIC)                                         # rint(xs), exact
    rem = xs - hi                                            # exact, |rem| <= 0.5
    lo = (rem * float(1 << (7 * nlo)) + MAGIC) - MAGIC
    dh = _digits(hi.astype(np.int64), 4)
    dl = _digits(lo.astype(np.int64), nlo)
    return np.stack(dh + dl)


def gram(B: np.ndarray, S: int, chunk_cols: int | None = None):
    """C ~= B B^T from the slices. Returns (C, info) with info = dict(max_abs_digit, max_abs_int32_accumulator)."""
    M, K = B.shape
    e = row_exponents(B)
    Q = slices(B, e, S)
    Kc = chunk_cols or (524288 // S)
    C = np.zeros((M, M))
    max_acc = 0
    for k0 in range(0, K, Kc):
        Qc = Q[:, :, k0:k0 + Kc]
        for u in range(S):                                   # anti-diagonal s + t = u, weight 2^-(12+7u)
            acc = np.zeros((M, M), dtype=np.int64)
            for s_ in range(u + 1):
                acc += Qc[s_] @ Qc[u - s_].T                 # exact integer GEMM
            max_acc = max(max_acc, int(np.abs(acc).max(initial=0)))
            C += acc.astype(np.float64) * 2.0 ** (-(12 + 7 * u))
    C = np.ldexp(C, e[:, None] + e[None, :])
    return C, dict(max_abs_digit=int(np.abs(Q).max(initial=0)), max_abs_int32_accumulator=max_acc, exponents=e, Q=Q)
