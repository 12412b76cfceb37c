// hiopLinSolverSymDense on device (B1): matrixChanged() / solve() semantics of
// src/LinAlg/hiopLinSolverSymDenseLapack.hpp:75-192 and the MAGMA twins hiopLinSolverSymDenseMagma.cpp:120-270, 324-476.
#include "hb_common.cuh"
#include "hb_dense.cuh"
#include <cstdlib>

struct hb_symdense
{
  hb_ctx* ctx = nullptr;
  int N = 0;
  double* M = nullptr;      // N x N row-major, upper triangle valid on entry, factor in place
  double* W = nullptr;      // panel scratch (lazy)
  double* xbuf = nullptr;   // staging for the *_host variants (lazy)
  size_t xbuf_cap = 0;      // doubles allocated in xbuf (per handle)
  double* Fpad = nullptr;   // odd N: copy of M with an even leading dimension (the large-N kernels use 16-byte accesses)
  double* F = nullptr;      // where the current factor lives (M or Fpad)
  long long ldf = 0;
  hb_big big;               // large-N path: streams, diagonal-block inverses, solve scratch
  bool big_solve = false;   // the current factor is solved with hb_big_solve
  // cluster Bunch-Kaufman (hb_bk_cluster.cu): permuted factor P A P^T = L D L^T
  bool bkc = false;
  double* dsub = nullptr;   // sub-diagonal of the 2x2 blocks of D
  int* perm = nullptr;      // gather order of the right-hand side
  int* bk_state = nullptr;  // device: k0, kb, info, -
  int* swaplog = nullptr;
  double* Wp = nullptr;     // W = L*D of the current panel
  int* ipiv = nullptr;
  int* info = nullptr;      // device: [0] info, [1..3] inertia
  int* info_host = nullptr; // pinned
  int mode = -1;
  bool factored = false;
  int n_neg = 0, n_null = 0, n_pos = 0;
};

namespace {
constexpr int BLOCKED_BK_MIN_N = 96; // below this the one-CTA unblocked DSYTF2 kernel is faster than panel + trailing launches
// Cholesky / no-pivot LDL^T: from this N on the look-ahead path of hb_dense_big.cu factors (below: cooperative / per-panel kernels);
// from BIG_SOLVE_MIN_N on the blocked multi-CTA solve replaces the one-CTA sweeps. HB_DENSE_BIG_MIN overrides the first (benchmarks).
int big_factor_min(int mode)
{
  static const int env = getenv("HB_DENSE_BIG_MIN") ? atoi(getenv("HB_DENSE_BIG_MIN")) : -1;
  if(env >= 0) return env;
  return mode == HB_FACT_CHOLESKY ? 1025 : 257;
}
constexpr int BIG_SOLVE_MIN_N = 257;
// Bunch-Kaufman: from this N on the cluster panel kernel factors (below: one-CTA DLASYF panels / DSYTF2). HB_BK_CLUSTER_MIN overrides.
int bk_cluster_min()
{
  static const int env = getenv("HB_BK_CLUSTER_MIN") ? atoi(getenv("HB_BK_CLUSTER_MIN")) : -1;
  return env >= 0 ? env : 385;
}
}

extern "C" int hb_symdense_create(hb_ctx* c, int N, hb_symdense** out)
{
  HB_REQUIRE(c && out && N >= 0, "hb_symdense_create: bad arguments");
  HB_CUDA(cudaSetDevice(c->device));
  hb_symdense* s = new hb_symdense;
  s->ctx = c;
  s->N = N;
  if(cudaMalloc(&s->M, sizeof(double) * (size_t)(N ? N : 1) * (N ? N : 1)) != cudaSuccess) {
    cudaGetLastError();
    delete s;
    return hb_fail(HB_ERR_ALLOC, "hb_symdense_create: cannot allocate the %s system matrix", "N x N");
  }
  HB_CUDA(cudaMalloc(&s->ipiv, sizeof(int) * (N + 1)));
  HB_CUDA(cudaMalloc(&s->info, sizeof(int) * 4));
  HB_CUDA(cudaMallocHost(&s->info_host, sizeof(int) * 4));
  HB_CUDA(cudaMemsetAsync(s->M, 0, sizeof(double) * (size_t)(N ? N : 1) * (N ? N : 1), c->stream));
  *out = s;
  return HB_OK;
}

extern "C" int hb_symdense_destroy(hb_symdense* s)
{
  if(!s) return HB_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  hb_big_release(&s->big);
  cudaFree(s->dsub); cudaFree(s->perm); cudaFree(s->bk_state); cudaFree(s->swaplog); cudaFree(s->Wp);
  cudaFree(s->M); cudaFree(s->W); cudaFree(s->xbuf); cudaFree(s->ipiv); cudaFree(s->info); cudaFree(s->Fpad);
  cudaFreeHost(s->info_host);
  delete s;
  return HB_OK;
}

extern "C" double* hb_symdense_matrix(hb_symdense* s) { return s ? s->M : nullptr; }

extern "C" int hb_symdense_matrix_changed(hb_symdense* s, int mode)
{
  HB_REQUIRE(s, "null handle");
  HB_REQUIRE(mode == HB_FACT_BUNCH_KAUFMAN || mode == HB_FACT_NOPIV || mode == HB_FACT_CHOLESKY, "hb_symdense_matrix_changed: bad mode");
  hb_ctx* c = s->ctx;
  const int N = s->N;
  s->mode = mode;
  s->factored = false;
  if(N == 0) { s->factored = true; s->n_neg = s->n_null = s->n_pos = 0; return 0; }
  HB_CUDA(cudaMemsetAsync(s->info, 0, sizeof(int) * 4, c->stream));
  s->F = s->M; s->ldf = N; s->big_solve = false; s->big.inv_valid = false; s->bkc = false;
  const bool bkc = (mode == HB_FACT_BUNCH_KAUFMAN && N >= bk_cluster_min() && hb_bkc_supported(c, N));
  const bool blocked_bk = (mode == HB_FACT_BUNCH_KAUFMAN && !bkc && N >= BLOCKED_BK_MIN_N);
  const bool big = (mode != HB_FACT_BUNCH_KAUFMAN && N >= big_factor_min(mode));
  if(big || bkc) {
    if(N & 1) { // even leading dimension for the 16-byte operand copies
      const long long ld = (N + 7) & ~7LL;
      if(!s->Fpad) {
        if(cudaMalloc(&s->Fpad, sizeof(double) * (size_t)ld * N) != cudaSuccess) { cudaGetLastError(); return hb_fail(HB_ERR_ALLOC, "hb_symdense_matrix_changed: cannot allocate the padded factor%s", ""); }
        HB_CUDA(cudaMemsetAsync(s->Fpad, 0, sizeof(double) * (size_t)ld * N, c->stream));
      }
      HB_CUDA(cudaMemcpy2DAsync(s->Fpad, sizeof(double) * ld, s->M, sizeof(double) * N, sizeof(double) * N, N, cudaMemcpyDeviceToDevice, c->stream));
      s->F = s->Fpad; s->ldf = ld;
    }
    if(big) {
      HB_CHECK(hb_big_factor(c, &s->big, N, s->F, s->ldf, mode == HB_FACT_NOPIV, s->info));
    } else {
      const long long ldw = (N + 7) & ~7LL;
      if(!s->dsub) {
        if(cudaMalloc(&s->dsub, sizeof(double) * (N + 2)) != cudaSuccess || cudaMalloc(&s->perm, sizeof(int) * (N + 2)) != cudaSuccess ||
           cudaMalloc(&s->bk_state, sizeof(int) * 4) != cudaSuccess || cudaMalloc(&s->swaplog, sizeof(int) * HB_BKC_SWAPLOG_INTS(N)) != cudaSuccess ||
           cudaMalloc(&s->Wp, sizeof(double) * HB_BKC_W_DOUBLES(ldw)) != cudaSuccess) {
          cudaGetLastError();
          return hb_fail(HB_ERR_ALLOC, "hb_symdense_matrix_changed: cannot allocate the Bunch-Kaufman scratch%s", "");
        }
      }
      HB_CHECK(hb_bkc_factor(c, &s->big, N, s->F, s->ldf, s->ipiv, s->dsub, s->perm, s->Wp, ldw, s->bk_state, s->swaplog, s->info));
      HB_CHECK(hb_big_block_inverses(c, &s->big, N, s->F, s->ldf, true));
      s->bkc = true;
    }
    s->big_solve = true;
  } else
  if((mode == HB_FACT_NOPIV || blocked_bk) && !s->W) {
    if(cudaMalloc(&s->W, sizeof(double) * (size_t)2 * 64 * N) != cudaSuccess) {
      cudaGetLastError();
      return hb_fail(HB_ERR_ALLOC, "hb_symdense_matrix_changed: cannot allocate panel scratch%s", "");
    }
  }
  if(big || bkc) {
  } else if(mode == HB_FACT_BUNCH_KAUFMAN) {
    if(blocked_bk) HB_CHECK(hb_dense_sytrf_blocked(c, N, s->M, N, s->ipiv, s->W, s->info));
    else HB_CHECK(hb_dense_sytf2(c, N, s->M, N, s->ipiv, s->info));
  } else {
    HB_CHECK(hb_dense_factor_blocked(c, N, s->M, N, mode == HB_FACT_NOPIV, s->W, s->info));
    if(N >= BIG_SOLVE_MIN_N) { // factor from the cooperative / per-panel kernels, solves through the blocked multi-CTA path
      HB_CHECK(hb_big_block_inverses(c, &s->big, N, s->M, N, mode == HB_FACT_NOPIV));
      s->big_solve = true;
    }
  }
  if(bkc) HB_CHECK(hb_bkc_inertia(c, N, s->F, s->ldf, s->ipiv, s->dsub, s->info + 1));
  else if(mode != HB_FACT_BUNCH_KAUFMAN) HB_CHECK(hb_bkc_inertia(c, N, s->F, s->ldf, nullptr, nullptr, s->info + 1)); // signs of the diagonal
  else HB_CHECK(hb_dense_inertia(c, N, s->F, (int)s->ldf, s->ipiv, mode, s->info + 1));
  HB_CUDA(cudaMemcpyAsync(s->info_host, s->info, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  s->n_neg = s->info_host[1]; s->n_null = s->info_host[2]; s->n_pos = s->info_host[3];
  if(s->info_host[0] != 0) return -1; // zero pivot / not SPD: "matrix is singular" (hiopLinSolverSymDenseLapack.hpp:109-116)
  s->factored = true;
  if(s->n_null > 0) return -1;        // :166
  return s->n_neg;
}

extern "C" int hb_symdense_inertia(hb_symdense* s, int* n_neg, int* n_null, int* n_pos)
{
  HB_REQUIRE(s, "null handle");
  if(n_neg) *n_neg = s->n_neg;
  if(n_null) *n_null = s->n_null;
  if(n_pos) *n_pos = s->n_pos;
  return HB_OK;
}

extern "C" int hb_symdense_solve(hb_symdense* s, double* x, int nrhs)
{
  HB_REQUIRE(s && nrhs >= 0, "hb_symdense_solve: bad arguments");
  if(s->N == 0 || nrhs == 0) return 1;
  HB_REQUIRE(x, "hb_symdense_solve: null rhs");
  if(!s->factored) return hb_fail(HB_ERR_STATE, "hb_symdense_solve: no valid factorization (call hb_symdense_matrix_changed)%s", "");
  hb_ctx* c = s->ctx;
  if(s->big_solve) {
    for(int r = 0; r < nrhs; r++)
      HB_CHECK(hb_big_solve(c, &s->big, s->N, s->F, s->ldf, s->bkc ? 2 : (s->mode == HB_FACT_CHOLESKY ? 0 : 1), s->ipiv, s->dsub, s->bkc ? s->perm : nullptr,
                            x + (size_t)r * s->N));
  } else if(s->mode == HB_FACT_BUNCH_KAUFMAN) {
    HB_CHECK(hb_dense_sytrs(c, s->N, s->M, s->N, s->ipiv, x, s->N, nrhs));
  } else {
    for(int r = 0; r < nrhs; r++) HB_CHECK(hb_dense_tri_solve(c, s->N, s->M, s->N, s->mode == HB_FACT_NOPIV, x + (size_t)r * s->N));
  }
  return 1;
}

extern "C" int hb_symdense_matrix_changed_host(hb_symdense* s, const double* M_host, int mode)
{
  HB_REQUIRE(s && (M_host || s->N == 0), "hb_symdense_matrix_changed_host: null matrix");
  if(s->N) HB_CUDA(cudaMemcpyAsync(s->M, M_host, sizeof(double) * (size_t)s->N * s->N, cudaMemcpyHostToDevice, s->ctx->stream));
  return hb_symdense_matrix_changed(s, mode);
}

extern "C" int hb_symdense_solve_host(hb_symdense* s, double* x_host, int nrhs)
{
  HB_REQUIRE(s && nrhs >= 0, "hb_symdense_solve_host: bad arguments");
  if(s->N == 0 || nrhs == 0) return 1;
  HB_REQUIRE(x_host, "hb_symdense_solve_host: null rhs");
  hb_ctx* c = s->ctx;
  const size_t need = (size_t)s->N * nrhs;
  if(!s->xbuf || s->xbuf_cap < need) {
    if(s->xbuf) { HB_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(s->xbuf); s->xbuf = nullptr; }
    if(cudaMalloc(&s->xbuf, sizeof(double) * need) != cudaSuccess) { cudaGetLastError(); return hb_fail(HB_ERR_ALLOC, "rhs staging allocation failed%s", ""); }
    s->xbuf_cap = need;
  }
  HB_CUDA(cudaMemcpyAsync(s->xbuf, x_host, sizeof(double) * need, cudaMemcpyHostToDevice, c->stream));
  int rc = hb_symdense_solve(s, s->xbuf, nrhs);
  if(rc != 1) return rc;
  HB_CUDA(cudaMemcpyAsync(x_host, s->xbuf, sizeof(double) * need, cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  return 1;
}

// diagnostics (tools/prof_diag.py): phase cycle counters of the 128 x 128 diagonal-block kernel on the leading block of M
extern "C" int hb_debug_diag128_profile(hb_symdense* s, int ldl, long long* prof_host8)
{
  HB_REQUIRE(s && prof_host8 && s->N >= 128 && (s->N & 1) == 0, "hb_debug_diag128_profile: needs an even N >= 128");
  return hb_big_diag_profile(s->ctx, &s->big, s->N, s->M, s->N, 0, ldl != 0, prof_host8);
}
extern "C" int hb_debug_bk_profile(hb_ctx* c, int on, long long* prof_host8)
{
  HB_REQUIRE(c, "null ctx");
  return hb_bkc_profile(c, on, prof_host8);
}
