// Internal API of hb_dense.cu (dense symmetric factorizations / solves on device).
#pragma once
#include "hb_common.cuh"

// Blocked LL^T (ldl=false) or no-pivot LDL^T (ldl=true) of the column-major-lower triangle of A (= row-major upper).
// Wpanel: 64*N doubles of scratch when ldl. info_dev: 0 ok, k>0 = breakdown at column k (1-based).
int hb_dense_factor_blocked(hb_ctx* c, int N, double* A, int lda, bool ldl, double* Wpanel, int* info_dev);
// Bunch-Kaufman: unblocked single-CTA kernel (DSYTF2 logic).
int hb_dense_sytf2(hb_ctx* c, int N, double* A, int lda, int* ipiv_dev, int* info_dev);
// Bunch-Kaufman: blocked (DLASYF panels + DMMA trailing updates); Wpanel: 2*64*N doubles of scratch.
int hb_dense_sytrf_blocked(hb_ctx* c, int N, double* A, int lda, int* ipiv_dev, double* Wpanel, int* info_dev);
int hb_dense_sytrs(hb_ctx* c, int N, const double* A, int lda, const int* ipiv_dev, double* B, int ldb, int nrhs);
int hb_dense_inertia(hb_ctx* c, int N, const double* A, int lda, const int* ipiv_dev, int mode, int* out3_dev);
int hb_dense_tri_solve(hb_ctx* c, int N, const double* F, int ldf, bool ldl, double* x);
int hb_dense_equilibrate(hb_ctx* c, int N, const double* Nfull, int ldn, double* F, int ldf, double* s);
int hb_dense_spd_solve_refine(hb_ctx* c, int N, const double* F, int ldf, const double* s, const double* Nref, int ldn, const double* rhs,
                              double* x, double* work2N, double tol, int max_refine, double* stats_dev);

// Cholesky that keeps the 16 x 16 diagonal inverses when the cooperative kernel runs (64 < N <= 2048), and the matching solve.
#define HB_CHOL_INV_DOUBLES(N) ((size_t)(((N) + 63) / 64) * (4 * 16 * 17))
int hb_dense_chol_with_inverses(hb_ctx* c, int N, double* A, int lda, int* info_dev, double* invd, bool* have_inv);
int hb_dense_spd_solve_refine2(hb_ctx* c, int N, const double* F, int ldf, const double* invd /* NULL: one-CTA solve */, const double* s,
                               const double* Nref, int ldn, const double* rhs, double* x, double* work2N2, double tol, int max_refine,
                               double* stats_dev);
