"""TEST INFRASTRUCTURE (never imported by hiop_b200/): numpy model of the panel algorithm of hb_bk_cluster.cu.

Bunch-Kaufman LDL^T (LAPACK DSYTF2 / DLASYF pivot rule, 'L' storage; role: hiopLinSolverSymDenseLapack::matrixChanged's DSYTRF,
src/LinAlg/hiopLinSolverSymDenseLapack.hpp:90-102) restated the way the device kernel organises it:

  * a panel of NB columns is kept in a "slab" (rows k0..N-1 x NB) that is updated RIGHT-LOOKING inside the panel, so the current
    column is always up to date and the pivot search needs no matrix-vector product;
  * the trailing matrix outside the panel stays NON-updated (w.r.t. this panel) in global memory and receives the symmetric
    interchanges as copies, exactly like DLASYF does; a pivot candidate column that lies outside the panel is updated on demand
    from the finished slab columns (upd = A(:,imax) - L * (D L(imax,:)^T));
  * L is written back (and W = L*D formed) once per panel; the interchanges are ALSO applied to all previous columns, i.e. the
    result is the fully permuted form  P A P^T = L D L^T  with one permutation vector (what dsytrs2/dsyconv use), not LAPACK's
    progressive row order.

Checked on the CPU (tests/test_cpu_bk_model.py) against scipy's dsytrf: identical pivot sequence (ipiv) and P A P^T = L D L^T."""
import numpy as np

ALPHA = (1.0 + np.sqrt(17.0)) / 8.0


def _dblock_apply(lrow, dinfo, kl):
    """v_c = (L D)(row, c) for the finished columns c < kl of the panel; dinfo[c] = ('1', d) or ('2a', d11, d21, d22) / ('2b',)"""
    v = np.zeros(kl)
    c = 0
    while c < kl:
        if dinfo[c][0] == "1":
            v[c] = lrow[c] * dinfo[c][1]
            c += 1
        else:
            _, d11, d21, d22 = dinfo[c]
            v[c] = lrow[c] * d11 + lrow[c + 1] * d21
            v[c + 1] = lrow[c] * d21 + lrow[c + 1] * d22
            c += 2
    return v


def factor(Afull, NB=32):
    """Afull: symmetric N x N. Returns (L, dblocks, perm, ipiv) with A[perm][:, perm] = L D L^T; ipiv in LAPACK's 1-based convention
    (positive = 1x1 with interchange row, negative pair = 2x2)."""
    N = Afull.shape[0]
    A = np.tril(Afull).astype(np.float64).copy()   # global matrix, lower part
    ipiv = np.zeros(N, dtype=np.int64)
    perm = np.arange(N)
    dsub = np.zeros(N)
    info = 0
    k0 = 0
    while k0 < N:
        last = N - k0 <= NB
        nbp = min(NB, N - k0)
        slab = A[k0:, k0:k0 + nbp].copy()           # slab[i - k0, c]; entries with row < col are garbage and never used
        dinfo = [None] * nbp
        swaps = []
        k = k0
        while k < N and (last or (k - k0) < NB - 1):
            kl = k - k0
            kstep = 1
            absakk = abs(slab[kl, kl])
            if k < N - 1:
                rel = int(np.argmax(np.abs(slab[kl + 1:, kl])))
                imax = k + 1 + rel
                colmax = abs(slab[imax - k0, kl])
            else:
                imax, colmax = k, 0.0
            kp = k
            ccol = None
            if max(absakk, colmax) == 0.0 or absakk != absakk:
                if info == 0:
                    info = k + 1
            elif absakk < ALPHA * colmax:
                # candidate column = row/column imax of the (updated) symmetric trailing matrix, entries for rows k..N-1
                ccol = np.zeros(N - k)
                il = imax - k0
                if imax < k0 + nbp:        # a panel column: the slab is current
                    for i in range(k, N):
                        ccol[i - k] = slab[il, i - k0] if i < imax else slab[i - k0, il]
                else:                      # outside the panel: global (non-updated) entries, updated on demand
                    v = _dblock_apply(slab[il, :kl], dinfo, kl)
                    for i in range(k, N):
                        if i < k0 + nbp:
                            ccol[i - k] = slab[il, i - k0]
                        else:
                            raw = A[imax, i] if i < imax else A[i, imax]
                            ccol[i - k] = raw - slab[i - k0, :kl] @ v
                tmp = np.abs(ccol).copy()
                tmp[imax - k] = -1.0
                rowmax = tmp.max()
                if absakk >= ALPHA * colmax * (colmax / rowmax):
                    kp = k
                elif abs(ccol[imax - k]) >= ALPHA * rowmax:
                    kp = imax
                else:
                    kp = imax
                    kstep = 2
            kk = k + kstep - 1
            kkl = kk - k0
            if kp != kk:
                kpl = kp - k0
                # --- global (non-updated) trailing matrix: DLASYF's copies of column kk into position kp ---
                A[kp, kp] = A[kk, kk]
                A[kp, kk + 1:kp] = A[kk + 1:kp, kk]
                A[kp + 1:, kp] = A[kp + 1:, kk]
                # --- slab (updated) ---
                old_col_kk = slab[:, kkl].copy()
                old_row_kk = slab[kkl, :].copy()
                old_row_kp = slab[kpl, :].copy()
                # rows kk <-> kp in the finished columns (and in column k for a 2x2 pivot)
                nfin = kl + (1 if kstep == 2 else 0)
                slab[kkl, :nfin] = old_row_kp[:nfin]
                slab[kpl, :nfin] = old_row_kk[:nfin]
                # row kp of the unfinished panel columns j in (kk, min(kp, k0+nbp)): T'(kp, j) = T(j, kk)
                for j in range(kk + 1, min(kp, k0 + nbp)):
                    slab[kpl, j - k0] = old_col_kk[j - k0]
                if kp < k0 + nbp:          # kp is a panel column: T'(i, kp) = T(i, kk) for i > kp, T'(kp,kp) = T(kk,kk)
                    slab[kpl + 1:, kpl] = old_col_kk[kpl + 1:]
                    slab[kpl, kpl] = old_col_kk[kkl]
                # new column kk = candidate column with positions kk and kp exchanged
                newcol = ccol.copy()
                newcol[kk - k], newcol[kp - k] = ccol[kp - k], ccol[kk - k]
                slab[kkl:, kkl] = newcol[kk - k:]
                swaps.append((kk, kp))
            elif kstep == 2:
                # kp == kk == k+1 == imax: the candidate column already is column k+1 of the slab
                pass
            if kstep == 1:
                d = slab[kl, kl]
                dinfo[kl] = ("1", d)
                if k < N - 1:
                    w = slab[kl + 1:, kl].copy()              # unscaled column
                    l = w / d if d != 0.0 else w
                    # rank-1 update of the unfinished panel columns c > kl (rows >= column)
                    for c in range(kl + 1, nbp):
                        slab[c:, c] -= l[c - kl - 1:] * w[c - kl - 1]
                    slab[kl + 1:, kl] = l
                ipiv[k] = kp + 1
            else:
                d11, d21, d22 = slab[kl, kl], slab[kl + 1, kl], slab[kl + 1, kl + 1]
                dinfo[kl] = ("2a", d11, d21, d22)
                dinfo[kl + 1] = ("2b",)
                dsub[k] = d21
                if k < N - 2:
                    w1 = slab[kl + 2:, kl].copy()
                    w2 = slab[kl + 2:, kl + 1].copy()
                    # LAPACK's scaled 2x2 inverse (dsytf2): d11' = d22/d21, d22' = d11/d21, t = 1/(d11' d22' - 1), s = t/d21
                    e11 = d22 / d21
                    e22 = d11 / d21
                    t = 1.0 / (e11 * e22 - 1.0)
                    s = t / d21
                    l1 = s * (e11 * w1 - w2)
                    l2 = s * (e22 * w2 - w1)
                    for c in range(kl + 2, nbp):
                        slab[c:, c] -= l1[c - kl - 2:] * w1[c - kl - 2] + l2[c - kl - 2:] * w2[c - kl - 2]
                    slab[kl + 2:, kl] = l1
                    slab[kl + 2:, kl + 1] = l2
                slab[kl + 1, kl] = 0.0                         # L(k+1,k) = 0; d21 kept in dsub
                ipiv[k] = -(kp + 1)
                ipiv[k + 1] = -(kp + 1)
            k += kstep
        kb = k - k0
        # ---- write back L (slab) for the factored columns; W = L*D for the rows below the panel; trailing update ----
        Lp = slab[:, :kb].copy()
        # put the D blocks on the diagonal (1x1: d; 2x2: d11, d22 on the diagonal, d21 in dsub) -- slab already has them there
        r0 = k0 + kb
        W = np.zeros((N - r0, kb))
        c = 0
        while c < kb:
            if dinfo[c][0] == "1":
                W[:, c] = Lp[kb:, c] * dinfo[c][1]
                c += 1
            else:
                _, d11, d21, d22 = dinfo[c]
                W[:, c] = Lp[kb:, c] * d11 + Lp[kb:, c + 1] * d21
                W[:, c + 1] = Lp[kb:, c] * d21 + Lp[kb:, c + 1] * d22
                c += 2
        for c in range(kb):
            A[k0 + c:, k0 + c] = Lp[c:, c]
        if r0 < N:
            upd = W @ Lp[kb:, :].T
            A[r0:, r0:] -= np.tril(upd)
        # ---- the panel's interchanges on all previous columns and on the permutation ----
        for (a, b) in swaps:
            if k0 > 0:
                A[[a, b], :k0] = A[[b, a], :k0]
            perm[[a, b]] = perm[[b, a]]
        k0 += kb
    L = np.tril(A, -1) + np.eye(N)
    dd = np.diag(A).copy()
    return L, dd, dsub, perm, ipiv, info


def dense_D(dd, dsub):
    N = dd.shape[0]
    D = np.diag(dd)
    for k in range(N - 1):
        if dsub[k] != 0.0:
            D[k + 1, k] = D[k, k + 1] = dsub[k]
    return D
