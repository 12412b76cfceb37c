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

// ---- large-N path (hb_dense_big.cu): blocked Cholesky / no-pivot LDL^T with look-ahead, 128 x 128 diagonal-block inverses, blocked solves ----
struct hb_big
{
  cudaStream_t panel_stream = nullptr;
  cudaEvent_t ev_panel = nullptr, ev_upd = nullptr, ev_upd2 = nullptr;
  double* InvAll = nullptr;         // ceil(N/128) inverses of the 128 x 128 diagonal triangles of the factor (column-major, zeros above)
  double* W[2] = {nullptr, nullptr}; // LDL^T: W = L*D of the current PAIR of panels, 256 p-major rows (double-buffered across the look-ahead)
  double* dinv = nullptr;
  double* partial = nullptr;        // solve: per-CTA partial products
  int* counter = nullptr;           // solve: ticket of the "last CTA finishes the step" pattern (self-resetting)
  double* xtmp = nullptr;           // permuted rhs (Bunch-Kaufman)
  int capN = 0;
  bool inv_valid = false;
};
int hb_big_init(hb_ctx* c, hb_big* b);
void hb_big_release(hb_big* b);
int hb_big_reserve(hb_ctx* c, hb_big* b, int N, bool need_w);
int hb_big_factor(hb_ctx* c, hb_big* b, int N, double* A, long long lda, bool ldl, int* info_dev);
int hb_big_diag_profile(hb_ctx* c, hb_big* b, int N, double* A, long long lda, int k0, bool ldl, long long* prof_host8);
int hb_big_trailing_from_state(hb_ctx* c, int N, double* A, long long lda, const double* W, long long ldw, const int* state_dev, int r0_min, cudaStream_t st);
int hb_big_block_inverses(hb_ctx* c, hb_big* b, int N, const double* F, long long ldf, bool unit);
int hb_big_solve(hb_ctx* c, hb_big* b, int N, const double* F, long long ldf, int dmode, const int* ipiv_dev, const double* dsub_dev, const int* perm_dev,
                 double* x);
// cluster Bunch-Kaufman (hb_bk_cluster.cu)
bool hb_bkc_supported(hb_ctx* c, int N);
int hb_bkc_factor(hb_ctx* c, hb_big* b, int N, double* A, long long lda, int* ipiv_dev, double* dsub_dev, int* perm_dev, double* Wp, long long ldw,
                  int* state_dev, int* swaplog_dev, int* info_dev);
int hb_bkc_inertia(hb_ctx* c, int N, const double* F, long long ldf, const int* ipiv_dev, const double* dsub_dev, int* out3_dev);
int hb_bkc_dsolve(hb_ctx* c, int N, const double* F, long long ldf, const int* ipiv_dev, const double* dsub_dev, double* x);
int hb_bkc_profile(hb_ctx* c, int on, long long* prof_host8);
#define HB_BKC_SWAPLOG_INTS(N) ((size_t)((N) / 7 + 4) * 132)
#define HB_BKC_W_DOUBLES(ldw) ((size_t)(ldw) * (64 + 128)) /* W = L*D of a panel (<= 64 columns) + staging rows of the interchange kernel */
