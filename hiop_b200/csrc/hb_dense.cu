// Dense symmetric factorizations and solves on device (FP64):
//   * blocked Cholesky LL^T and no-pivot LDL^T  (DPOTRF / magma_dsytrf_nopiv_gpu roles)
//   * Bunch-Kaufman LDL^T with the LAPACK pivoting rule (DSYTRF / magma_dsytrf_gpu roles) + DSYTRS
//   * inertia from the 1x1 / 2x2 pivots with the reference's dsidi rule and +-1e-14 thresholds
//     (src/LinAlg/hiopLinSolverSymDenseLapack.hpp:127-167)
//   * equilibrated SPD solve with device-side residual-driven refinement (hiopKKTLinSysLowRank::solveWithRefin,
//     src/Optimization/hiopKKTLinSys.cpp:1192-1350)
//
// Storage convention of the whole file: the system matrix is N x N ROW-major with its UPPER triangle valid
// (hiopKKTLinSysMDS.cpp:196-206). Read column-major that is the LOWER triangle, which is how LAPACK sees it
// (uplo='L', hiopLinSolverSymDenseLapack.hpp:84) and how the kernels index it:  Lc(i,j) = A[j*lda + i], i >= j.
// Column j of the factor is therefore contiguous in memory -> coalesced along i.
#include "hb_common.cuh"
#include "hb_dense.cuh"

namespace {

#define LC(A, lda, i, j) (A)[(size_t)(j) * (lda) + (i)]

constexpr int NB = 64;             // panel width of the blocked factorizations
constexpr int PANEL_THREADS = 256; // one row of the slab per thread

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------------------
// Panel kernel: every CTA factors the nb x nb diagonal block in shared memory (redundantly -- it is 64^3/3 flops),
// CTA 0 writes it back, and each CTA then computes its 256-row slab of L21 by forward substitution, one row per
// thread. LDL variant also emits W = L21*D (needed by the trailing update) into Wout (nb rows of length ldw).
// ---------------------------------------------------------------------------------------------------------
template <bool LDL>
__global__ void __launch_bounds__(PANEL_THREADS)
k_panel(double* __restrict__ A, int lda, int N, int k0, int nb, double* __restrict__ Wout, int ldw, int* __restrict__ info)
{
  __shared__ double D[NB][NB + 1];
  __shared__ double dinv[NB];
  const int tid = threadIdx.x;
  // load the lower part of the diagonal block; pad to identity beyond nb
  {
    double v[NB * NB / PANEL_THREADS];
#pragma unroll
    for(int q = 0; q < NB * NB / PANEL_THREADS; q++) { // all loads first (independent), then the stores
      const int e = tid + q * PANEL_THREADS;
      const int j = e / NB, i = e % NB;
      v[q] = (i == j) ? 1.0 : 0.0;
      if(i < nb && j < nb && i >= j) v[q] = LC(A, lda, k0 + i, k0 + j);
    }
#pragma unroll
    for(int q = 0; q < NB * NB / PANEL_THREADS; q++) {
      const int e = tid + q * PANEL_THREADS;
      D[e / NB][e % NB] = v[q]; // D[j][i] holds element (i,j)
    }
  }
  __syncthreads();
  if(!LDL) {
    // Blocked right-looking Cholesky of the 64x64 diagonal block in shared memory, 16-wide sub-panels, 3 barriers per
    // sub-panel (12 in total instead of 2-3 per column):
    //   (1) warp 0 factors the 16x16 diagonal sub-block in registers (lane r holds row r) with shuffles,
    //   (2) one thread per row below solves its 16 entries against it,
    //   (3) all threads apply the rank-16 update to the remaining lower triangle.
    // D beyond nb is the identity, so a partial last panel needs no special casing.
    const int lane = tid & 31, warp = tid >> 5;
    for(int kb = 0; kb < NB; kb += 16) {
      if(warp == 0) {
        double a[16];
#pragma unroll
        for(int c = 0; c < 16; c++) a[c] = (lane < 16 && c <= lane) ? D[kb + c][kb + lane] : 0.0;
        // spelled out per column: the 16 x 15 nest does not get fully unrolled otherwise and a[] lands in local memory
#define PANEL_CHOL_COL(j)                                                                      \
  {                                                                                            \
    const double d = __shfl_sync(0xffffffffu, a[j], j);                                        \
    if(!(d > 0.0) && lane == 0 && blockIdx.x == 0) atomicCAS(info, 0, k0 + kb + j + 1);       \
    const double l = sqrt(d);                                                                  \
    if(lane == j) a[j] = l;                                                                    \
    else if(lane > j) a[j] = a[j] / l;                                                         \
    _Pragma("unroll") for(int c = j + 1; c < 16; c++) {                                        \
      const double lc = __shfl_sync(0xffffffffu, a[j], c);                                     \
      if(lane >= c) a[c] -= a[j] * lc;                                                         \
    }                                                                                          \
  }
        PANEL_CHOL_COL(0) PANEL_CHOL_COL(1) PANEL_CHOL_COL(2) PANEL_CHOL_COL(3) PANEL_CHOL_COL(4) PANEL_CHOL_COL(5) PANEL_CHOL_COL(6)
        PANEL_CHOL_COL(7) PANEL_CHOL_COL(8) PANEL_CHOL_COL(9) PANEL_CHOL_COL(10) PANEL_CHOL_COL(11) PANEL_CHOL_COL(12) PANEL_CHOL_COL(13)
        PANEL_CHOL_COL(14) PANEL_CHOL_COL(15)
#undef PANEL_CHOL_COL
#pragma unroll
        for(int c = 0; c < 16; c++)
          if(lane < 16 && c <= lane) D[kb + c][kb + lane] = a[c];
      }
      __syncthreads();
      const int below = NB - kb - 16;
      if(tid < below) {
        const int r = kb + 16 + tid;
        double x[16];
#pragma unroll
        for(int c = 0; c < 16; c++) x[c] = D[kb + c][r];
#pragma unroll
        for(int j = 0; j < 16; j++) {
          x[j] /= D[kb + j][kb + j];
          const double xj = x[j];
#pragma unroll
          for(int q = j + 1; q < 16; q++) x[q] -= xj * D[kb + j][kb + q];
        }
#pragma unroll
        for(int c = 0; c < 16; c++) D[kb + c][r] = x[c];
      }
      __syncthreads();
      for(int e = tid; e < below * below; e += PANEL_THREADS) {
        const int c = kb + 16 + e / below, i = kb + 16 + e % below;
        if(i >= c) {
          double s = 0.0;
#pragma unroll
          for(int p = 0; p < 16; p++) s += D[kb + p][i] * D[kb + p][c];
          D[c][i] -= s;
        }
      }
      __syncthreads();
    }
  } else {
  for(int j = 0; j < nb; j++) {
    const double ajj = D[j][j];
    {
      if(ajj == 0.0 || ajj != ajj) {
        if(tid == 0 && blockIdx.x == 0) atomicCAS(info, 0, k0 + j + 1);
      }
      const double r = 1.0 / ajj;
      const int rem = nb - j - 1;
      __syncthreads();
      // trailing uses w = column j (unscaled) and l = w / d
      for(int e = tid; e < rem * rem; e += PANEL_THREADS) {
        const int c = j + 1 + e / rem, i = j + 1 + e % rem;
        if(i >= c) D[c][i] -= D[j][i] * r * D[j][c];
      }
      __syncthreads();
      for(int i = j + 1 + tid; i < nb; i += PANEL_THREADS) D[j][i] *= r;
      if(tid == 0) dinv[j] = r;
      __syncthreads();
    }
  }
  }
  if(blockIdx.x == 0) {
    for(int e = tid; e < nb * nb; e += PANEL_THREADS) {
      const int j = e / nb, i = e % nb;
      if(i >= j) LC(A, lda, k0 + i, k0 + j) = D[j][i];
    }
  }
  // slab rows
  const int i = k0 + nb + blockIdx.x * PANEL_THREADS + tid;
  if(i < N) {
    double x[NB];
#pragma unroll
    for(int j = 0; j < NB; j++) x[j] = j < nb ? LC(A, lda, i, k0 + j) : 0.0;
    // right-looking forward substitution: once x[j] is final, all later entries are updated independently (ILP 63..1
    // instead of a dependent chain per entry)
#pragma unroll
    for(int j = 0; j < NB; j++) {
      if(!LDL) x[j] /= D[j][j];
      const double xj = x[j];
#pragma unroll
      for(int q = j + 1; q < NB; q++) x[q] -= xj * D[j][q]; // element (q,j) of L11
    }
    if(LDL) {
#pragma unroll
      for(int j = 0; j < NB; j++)
        if(j < nb) {
          Wout[(size_t)j * ldw + i] = x[j];
          LC(A, lda, i, k0 + j) = x[j] * dinv[j];
        }
    } else {
#pragma unroll
      for(int j = 0; j < NB; j++)
        if(j < nb) LC(A, lda, i, k0 + j) = x[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Trailing update on the DMMA pipe: for i >= j >= r0:  Lc(i,j) -= sum_p P[p][i] * Q[p][j],  p < kb.
// P, Q are kb "row-contiguous" panels: P[p][i] = Pbase[p*ldp + i] (for Cholesky P = Q = the factor panel rows,
// for LDL^T P = W = L*D, Q = L). 64x64 output tiles, 4 warps (2x2), warp tile 32x32.
// ---------------------------------------------------------------------------------------------------------
constexpr int TT = 64;
constexpr int TLD = TT + 4; // padded: fragment reads (p = lane%4, i = lane/4) are bank-conflict free
__global__ void __launch_bounds__(128)
k_trailing(double* __restrict__ A, int lda, int N, int r0, const double* __restrict__ P, long long ldp, const double* __restrict__ Q,
           long long ldq, int kb)
{
  extern __shared__ __align__(16) unsigned char trailing_smem[];
  double (*sP)[TLD] = reinterpret_cast<double (*)[TLD]>(trailing_smem);
  double (*sQ)[TLD] = sP + NB;
  const int nt = (N - r0 + TT - 1) / TT;
  // linear tile id -> (ti >= tj)
  int t = blockIdx.x, ti = 0;
  while(t >= ti + 1) { t -= ti + 1; ti++; }
  const int tj = t;
  (void)nt;
  const int i0 = r0 + ti * TT, j0 = r0 + tj * TT;
  const int tid = threadIdx.x;
  {
    // all 8-byte copies of the two operand tiles are issued back to back (one L2 latency for the whole tile instead of
    // one per loop iteration); columns beyond N and the K padding are zero-filled through the src-size operand
    const int kpad = ((kb + 3) / 4) * 4;
    for(int e = tid; e < kpad * TT; e += 128) {
      const int p = e / TT, c = e % TT;
      const bool vp = (p < kb) && (i0 + c < N), vq = (p < kb) && (j0 + c < N);
      const unsigned sp = (unsigned)__cvta_generic_to_shared(&sP[p][c]), sq = (unsigned)__cvta_generic_to_shared(&sQ[p][c]);
      const double* gp = vp ? P + (size_t)p * ldp + i0 + c : P;
      const double* gq = vq ? Q + (size_t)p * ldq + j0 + c : Q;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sp), "l"(gp), "r"(vp ? 8 : 0));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sq), "l"(gq), "r"(vq ? 8 : 0));
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int wi = warp & 1, wj = warp >> 1;
  const int g = lane >> 2, t4 = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for(int a = 0; a < 4; a++)
#pragma unroll
    for(int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
  const int ksteps = (kb + 3) / 4;
  for(int kk = 0; kk < ksteps; kk++) {
    double af[4], bf[4];
#pragma unroll
    for(int a = 0; a < 4; a++) af[a] = sP[kk * 4 + t4][wi * 32 + a * 8 + g];
#pragma unroll
    for(int b = 0; b < 4; b++) bf[b] = sQ[kk * 4 + t4][wj * 32 + b * 8 + g];
#pragma unroll
    for(int a = 0; a < 4; a++)
#pragma unroll
      for(int b = 0; b < 4; b++) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
  }
  // epilogue: all loads of the C tile first, then all stores (a load->sub->store chain per element would serialise on
  // L2 latency because the compiler must assume the stores alias the following loads)
  double cv[4][4][2];
#pragma unroll
  for(int a = 0; a < 4; a++) {
    const int i = i0 + wi * 32 + a * 8 + g;
#pragma unroll
    for(int b = 0; b < 4; b++)
#pragma unroll
      for(int h = 0; h < 2; h++) {
        const int j = j0 + wj * 32 + b * 8 + t4 * 2 + h;
        cv[a][b][h] = (i < N && j < N && i >= j) ? LC(A, lda, i, j) : 0.0;
      }
  }
#pragma unroll
  for(int a = 0; a < 4; a++) {
    const int i = i0 + wi * 32 + a * 8 + g;
#pragma unroll
    for(int b = 0; b < 4; b++)
#pragma unroll
      for(int h = 0; h < 2; h++) {
        const int j = j0 + wj * 32 + b * 8 + t4 * 2 + h;
        if(i < N && j < N && i >= j) LC(A, lda, i, j) = cv[a][b][h] - acc[a][b][h];
      }
  }
}

// ---------------------------------------------------------------------------------------------------------
// CTA-wide triangular solves with a column-major-lower factor (1024 threads). x lives in global memory.
// unit_diag: LDL^T factors (L has an implicit unit diagonal).
// ---------------------------------------------------------------------------------------------------------
constexpr int SOLVE_THREADS = 1024;

// sd: 32 x 33 doubles of shared memory holding the current diagonal block (sd[c][r] = L(j0+r, j0+c)): one coalesced
// load instead of 32 dependent L2 round trips inside the sequential part.
__device__ void dev_forward(const double* __restrict__ A, int lda, int N, double* x, bool unit_diag, double (*sd)[33])
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for(int j0 = 0; j0 < N; j0 += 32) {
    const int nb = min(32, N - j0);
    {
      const int c = tid >> 5, r = tid & 31; // 1024 threads = one 32x32 block
      if(c < nb && r < nb && r >= c) sd[c][r] = LC(A, lda, j0 + r, j0 + c);
    }
    __syncthreads();
    if(warp == 0) {
      double b = lane < nb ? x[j0 + lane] : 0.0;
      for(int c = 0; c < nb; c++) {
        double yc = __shfl_sync(0xffffffffu, b, c);
        if(!unit_diag) yc /= sd[c][c];
        if(lane == c) b = yc;
        if(lane > c && lane < nb) b -= sd[c][lane] * yc;
      }
      if(lane < nb) x[j0 + lane] = b;
    }
    __syncthreads();
    for(int i = j0 + nb + tid; i < N; i += SOLVE_THREADS) {
      double s = x[i];
#pragma unroll 8
      for(int c = 0; c < nb; c++) s -= LC(A, lda, i, j0 + c) * x[j0 + c];
      x[i] = s;
    }
    __syncthreads();
  }
}

__device__ void dev_backward(const double* __restrict__ A, int lda, int N, double* x, bool unit_diag, double* sm32 /* 32 doubles */,
                             double (*sd)[33])
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nblk = (N + 31) / 32;
  for(int bi = nblk - 1; bi >= 0; bi--) {
    const int j0 = bi * 32;
    const int nb = min(32, N - j0);
    {
      const int c = tid >> 5, r = tid & 31;
      if(c < nb && r < nb && r >= c) sd[c][r] = LC(A, lda, j0 + r, j0 + c);
    }
    // each warp: dot of column (j0+warp) below the block with the already solved tail of x
    if(warp < nb) {
      double s = 0.0;
#pragma unroll 8
      for(int i = j0 + nb + lane; i < N; i += 32) s += LC(A, lda, i, j0 + warp) * x[i];
      s = hb_warp_sum(s);
      if(lane == 0) sm32[warp] = s;
    }
    __syncthreads();
    if(warp == 0) {
      double b = lane < nb ? x[j0 + lane] - sm32[lane] : 0.0;
      for(int c = nb - 1; c >= 0; c--) {
        // x_c = (b_c - sum_{t>c} L(t,c) x_t) / L(c,c)
        double part = (lane > c && lane < nb) ? sd[c][lane] * b : 0.0;
        part = hb_warp_sum(part);
        if(lane == c) {
          b = b - part;
          if(!unit_diag) b /= sd[c][c];
        }
      }
      if(lane < nb) x[j0 + lane] = b;
    }
    __syncthreads();
  }
}

// One-CTA SPD solve with equilibration scaling s and device-side refinement against the unscaled matrix Nref
// (full symmetric storage, row-major, ld = ldn). stats: [0]=#refinements, [1]=last residual inf-norm, [2]=info.
__global__ void __launch_bounds__(SOLVE_THREADS)
k_spd_solve_refine(const double* __restrict__ F, int ldf, int N, const double* __restrict__ s, const double* __restrict__ Nref, int ldn,
                   const double* __restrict__ rhs, double* __restrict__ x, double* __restrict__ work /* 2N */, double tol, int max_refine,
                   double* __restrict__ stats)
{
  __shared__ double sm[32];
  __shared__ double sd[32][33];
  __shared__ double s_nrm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* r = work;       // residual / correction
  double* z = work + N;   // scaled rhs
  for(int i = tid; i < N; i += SOLVE_THREADS) z[i] = rhs[i] * s[i];
  __syncthreads();
  dev_forward(F, ldf, N, z, false, sd);
  dev_backward(F, ldf, N, z, false, sm, sd);
  for(int i = tid; i < N; i += SOLVE_THREADS) x[i] = z[i] * s[i];
  __syncthreads();
  int nref = 0;
  double nrm = 0.0;
  while(true) {
    // r = rhs - Nref*x : one warp per row (rows are contiguous)
    double wmax = 0.0;
    for(int i = warp; i < N; i += SOLVE_THREADS / 32) {
      double acc = 0.0;
      const double* row = Nref + (size_t)i * ldn;
      for(int j = lane; j < N; j += 32) acc += row[j] * x[j];
      acc = hb_warp_sum(acc);
      const double ri = rhs[i] - acc;
      if(lane == 0) r[i] = ri;
      wmax = fmax(wmax, fabs(ri));
    }
    __syncthreads();
    if(lane == 0) sm[warp] = wmax;
    __syncthreads();
    if(tid == 0) {
      double m = 0.0;
      for(int w = 0; w < SOLVE_THREADS / 32; w++) m = fmax(m, sm[w]);
      s_nrm = m;
    }
    __syncthreads();
    nrm = s_nrm;
    if(!(nrm >= tol) || nref >= max_refine) break; // also leaves on NaN
    for(int i = tid; i < N; i += SOLVE_THREADS) r[i] *= s[i];
    __syncthreads();
    dev_forward(F, ldf, N, r, false, sd);
    dev_backward(F, ldf, N, r, false, sm, sd);
    for(int i = tid; i < N; i += SOLVE_THREADS) x[i] += r[i] * s[i];
    __syncthreads();
    nref++;
  }
  if(tid == 0) {
    stats[0] = (double)nref;
    stats[1] = nrm;
  }
}

// s_i = 1/sqrt(A_ii); F(i,j) = s_i A(i,j) s_j on the column-major-lower triangle (equilibration of DPOSVX('E'):
// always applied here -- a symmetric diagonal scaling never changes the exact solution).
__global__ void k_equilibrate(const double* __restrict__ Nfull, int ldn, int N, double* __restrict__ F, int ldf, double* __restrict__ s)
{
  const long long total = (long long)N * N;
  for(long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(e / N), i = (int)(e % N);
    const double sj = 1.0 / sqrt(Nfull[(size_t)j * ldn + j]);
    if(i == j) s[j] = sj;
    if(i >= j) {
      const double si = 1.0 / sqrt(Nfull[(size_t)i * ldn + i]);
      LC(F, ldf, i, j) = si * Nfull[(size_t)j * ldn + i] * sj;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Unblocked Bunch-Kaufman (DSYTF2 'L' logic, bit-compatible pivot choices with LAPACK) in one CTA.
// Used for the 2l x 2l matrix V of the compact BFGS inverse and for small KKT systems; the blocked variant for
// large N is k_lasyf_panel + k_trailing below.
// ---------------------------------------------------------------------------------------------------------
constexpr int BK_THREADS = 1024;
#define BK_ALPHA 0.6403882032022076 /* (1+sqrt(17))/8 */

struct ArgMax
{
  double v;
  int i;
};
__device__ __forceinline__ ArgMax argmax_comb(ArgMax a, ArgMax b)
{
  // IDAMAX semantics: first index of the maximum absolute value
  if(b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ ArgMax block_argmax(ArgMax a, ArgMax* sm /* 32 */)
{
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = argmax_comb(a, b);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if(lane == 0) sm[warp] = a;
  __syncthreads();
  ArgMax r = sm[0];
  const int nw = blockDim.x >> 5;
  for(int w = 1; w < nw; w++) r = argmax_comb(r, sm[w]);
  return r; // every thread gets the same answer
}

__global__ void __launch_bounds__(BK_THREADS)
k_sytf2(double* __restrict__ A, int lda, int N, int* __restrict__ ipiv, int* __restrict__ info)
{
  __shared__ ArgMax sm[32];
  const int tid = threadIdx.x;
  const int big = 0x7fffffff;
  int k = 0;
  int linfo = 0;
  while(k < N) {
    int kstep = 1, kp = k;
    const double absakk = fabs(LC(A, lda, k, k));
    int imax = k;
    double colmax = 0.0;
    if(k < N - 1) {
      ArgMax a{-1.0, big};
      for(int i = k + 1 + tid; i < N; i += BK_THREADS) a = argmax_comb(a, ArgMax{fabs(LC(A, lda, i, k)), i});
      a = block_argmax(a, sm);
      imax = a.i;
      colmax = a.v;
    }
    if(fmax(absakk, colmax) == 0.0 || absakk != absakk) {
      if(linfo == 0) linfo = k + 1;
      kp = k;
    } else {
      if(absakk >= BK_ALPHA * colmax) {
        kp = k;
      } else {
        ArgMax a{-1.0, big};
        for(int j = k + tid; j < imax; j += BK_THREADS) a = argmax_comb(a, ArgMax{fabs(LC(A, lda, imax, j)), j});
        for(int i = imax + 1 + tid; i < N; i += BK_THREADS) a = argmax_comb(a, ArgMax{fabs(LC(A, lda, i, imax)), i});
        a = block_argmax(a, sm);
        const double rowmax = a.v;
        if(absakk >= BK_ALPHA * colmax * (colmax / rowmax)) kp = k;
        else if(fabs(LC(A, lda, imax, imax)) >= BK_ALPHA * rowmax) kp = imax;
        else { kp = imax; kstep = 2; }
      }
      const int kk = k + kstep - 1;
      __syncthreads();
      if(kp != kk) {
        for(int i = kp + 1 + tid; i < N; i += BK_THREADS) {
          const double t = LC(A, lda, i, kk);
          LC(A, lda, i, kk) = LC(A, lda, i, kp);
          LC(A, lda, i, kp) = t;
        }
        for(int i = kk + 1 + tid; i < kp; i += BK_THREADS) {
          const double t = LC(A, lda, i, kk);
          LC(A, lda, i, kk) = LC(A, lda, kp, i);
          LC(A, lda, kp, i) = t;
        }
        if(tid == 0) {
          double t = LC(A, lda, kk, kk);
          LC(A, lda, kk, kk) = LC(A, lda, kp, kp);
          LC(A, lda, kp, kp) = t;
          if(kstep == 2) {
            t = LC(A, lda, k + 1, k);
            LC(A, lda, k + 1, k) = LC(A, lda, kp, k);
            LC(A, lda, kp, k) = t;
          }
        }
        __syncthreads();
      }
      if(kstep == 1) {
        if(k < N - 1) {
          const double d11 = 1.0 / LC(A, lda, k, k);
          const int rem = N - k - 1;
          // A(i,j) -= d11 * x_i * x_j, k < j <= i
          for(long long e = tid; e < (long long)rem * rem; e += BK_THREADS) {
            const int j = k + 1 + (int)(e / rem), i = k + 1 + (int)(e % rem);
            if(i >= j) LC(A, lda, i, j) -= d11 * LC(A, lda, i, k) * LC(A, lda, j, k);
          }
          __syncthreads();
          for(int i = k + 1 + tid; i < N; i += BK_THREADS) LC(A, lda, i, k) *= d11;
          __syncthreads();
        }
      } else {
        if(k < N - 2) {
          double d21 = LC(A, lda, k + 1, k);
          const double d11 = LC(A, lda, k + 1, k + 1) / d21;
          const double d22 = LC(A, lda, k, k) / d21;
          const double t = 1.0 / (d11 * d22 - 1.0);
          d21 = t / d21;
          const int rem = N - k - 2;
          for(long long e = tid; e < (long long)rem * rem; e += BK_THREADS) {
            const int j = k + 2 + (int)(e / rem), i = k + 2 + (int)(e % rem);
            if(i >= j) {
              const double wk = d21 * (d11 * LC(A, lda, j, k) - LC(A, lda, j, k + 1));
              const double wkp1 = d21 * (d22 * LC(A, lda, j, k + 1) - LC(A, lda, j, k));
              LC(A, lda, i, j) = LC(A, lda, i, j) - LC(A, lda, i, k) * wk - LC(A, lda, i, k + 1) * wkp1;
            }
          }
          __syncthreads();
          for(int j = k + 2 + tid; j < N; j += BK_THREADS) {
            const double ajk = LC(A, lda, j, k), ajk1 = LC(A, lda, j, k + 1);
            LC(A, lda, j, k) = d21 * (d11 * ajk - ajk1);
            LC(A, lda, j, k + 1) = d21 * (d22 * ajk1 - ajk);
          }
          __syncthreads();
        }
      }
    }
    if(tid == 0) {
      if(kstep == 1) ipiv[k] = kp + 1;
      else { ipiv[k] = -(kp + 1); ipiv[k + 1] = -(kp + 1); }
    }
    k += kstep;
    __syncthreads();
  }
  if(tid == 0) *info = linfo;
}

// DSYTRS 'L': one thread per right-hand side (rhs r = B + r*ldb, contiguous N doubles). Used for V^{-1}[S1^T;Y1^T]
// (many rhs, tiny N) and for single-rhs solves with small N.
__global__ void k_sytrs_per_rhs(const double* __restrict__ A, int lda, int N, const int* __restrict__ ipiv, double* __restrict__ B, int ldb,
                                int nrhs)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if(r >= nrhs) return;
  double* b = B + (size_t)r * ldb;
  int k = 0;
  while(k < N) {
    if(ipiv[k] > 0) {
      const int kp = ipiv[k] - 1;
      if(kp != k) { const double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      const double bk = b[k];
      for(int i = k + 1; i < N; i++) b[i] -= LC(A, lda, i, k) * bk;
      b[k] = bk / LC(A, lda, k, k);
      k += 1;
    } else {
      const int kp = -ipiv[k] - 1;
      if(kp != k + 1) { const double t = b[k + 1]; b[k + 1] = b[kp]; b[kp] = t; }
      const double bk0 = b[k], bk1 = b[k + 1];
      for(int i = k + 2; i < N; i++) b[i] -= LC(A, lda, i, k) * bk0 + LC(A, lda, i, k + 1) * bk1;
      const double akm1k = LC(A, lda, k + 1, k);
      const double akm1 = LC(A, lda, k, k) / akm1k;
      const double ak = LC(A, lda, k + 1, k + 1) / akm1k;
      const double denom = akm1 * ak - 1.0;
      const double bkm1 = bk0 / akm1k, bkk = bk1 / akm1k;
      b[k] = (ak * bkm1 - bkk) / denom;
      b[k + 1] = (akm1 * bkk - bkm1) / denom;
      k += 2;
    }
  }
  k = N - 1;
  while(k >= 0) {
    if(ipiv[k] > 0) {
      double s = b[k];
      for(int i = k + 1; i < N; i++) s -= LC(A, lda, i, k) * b[i];
      b[k] = s;
      const int kp = ipiv[k] - 1;
      if(kp != k) { const double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      k -= 1;
    } else {
      double s0 = b[k], s1 = b[k - 1];
      for(int i = k + 1; i < N; i++) {
        s0 -= LC(A, lda, i, k) * b[i];
        s1 -= LC(A, lda, i, k - 1) * b[i];
      }
      b[k] = s0;
      b[k - 1] = s1;
      const int kp = -ipiv[k] - 1;
      if(kp != k) { const double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      k -= 2;
    }
  }
}

// DSYTRS 'L' for ONE right-hand side with the whole CTA cooperating on each column sweep (large N).
__global__ void __launch_bounds__(SOLVE_THREADS)
k_sytrs_cta(const double* __restrict__ A, int lda, int N, const int* __restrict__ ipiv, double* __restrict__ b)
{
  __shared__ double sm[64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int k = 0;
  while(k < N) {
    const int pv = ipiv[k];
    if(pv > 0) {
      const int kp = pv - 1;
      if(tid == 0 && kp != k) { const double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      __syncthreads();
      const double bk = b[k];
      for(int i = k + 1 + tid; i < N; i += SOLVE_THREADS) b[i] -= LC(A, lda, i, k) * bk;
      __syncthreads();
      if(tid == 0) b[k] = bk / LC(A, lda, k, k);
      k += 1;
    } else {
      const int kp = -pv - 1;
      if(tid == 0 && kp != k + 1) { const double t = b[k + 1]; b[k + 1] = b[kp]; b[kp] = t; }
      __syncthreads();
      const double bk0 = b[k], bk1 = b[k + 1];
      for(int i = k + 2 + tid; i < N; i += SOLVE_THREADS) b[i] -= LC(A, lda, i, k) * bk0 + LC(A, lda, i, k + 1) * bk1;
      __syncthreads();
      if(tid == 0) {
        const double akm1k = LC(A, lda, k + 1, k);
        const double akm1 = LC(A, lda, k, k) / akm1k;
        const double ak = LC(A, lda, k + 1, k + 1) / akm1k;
        const double denom = akm1 * ak - 1.0;
        const double bkm1 = bk0 / akm1k, bkk = bk1 / akm1k;
        b[k] = (ak * bkm1 - bkk) / denom;
        b[k + 1] = (akm1 * bkk - bkm1) / denom;
      }
      k += 2;
    }
    __syncthreads();
  }
  k = N - 1;
  while(k >= 0) {
    const int pv = ipiv[k];
    const int ncol = pv > 0 ? 1 : 2;
    double s0 = 0.0, s1 = 0.0;
    for(int i = k + 1 + tid; i < N; i += SOLVE_THREADS) {
      const double bi = b[i];
      s0 += LC(A, lda, i, k) * bi;
      if(ncol == 2) s1 += LC(A, lda, i, k - 1) * bi;
    }
    s0 = hb_warp_sum(s0);
    s1 = hb_warp_sum(s1);
    __syncthreads();
    if(lane == 0) { sm[warp] = s0; sm[32 + warp] = s1; }
    __syncthreads();
    if(tid == 0) {
      double t0 = 0.0, t1 = 0.0;
      for(int w = 0; w < SOLVE_THREADS / 32; w++) { t0 += sm[w]; t1 += sm[32 + w]; }
      b[k] -= t0;
      if(ncol == 2) b[k - 1] -= t1;
      const int kp = (pv > 0 ? pv : -pv) - 1;
      if(kp != k) { const double t = b[k]; b[k] = b[kp]; b[kp] = t; }
    }
    __syncthreads();
    k -= ncol;
  }
}

// Inertia sweep: BK factor -> LINPACK dsidi rule; no-pivot LDL^T / Cholesky -> signs of the diagonal.
// out = {neg, null, pos}
__global__ void __launch_bounds__(1024)
k_inertia(const double* __restrict__ A, int lda, int N, const int* __restrict__ ipiv, int mode, int* __restrict__ out, double* __restrict__ scratch)
{
  // scratch: 2N doubles (diagonal, sub-diagonal), gathered by all threads (strided loads in parallel)
  double* dg = scratch;
  double* sub = scratch + N;
  for(int k = threadIdx.x; k < N; k += blockDim.x) {
    dg[k] = LC(A, lda, k, k);
    sub[k] = (k + 1 < N) ? LC(A, lda, k + 1, k) : 0.0;
  }
  __syncthreads();
  if(threadIdx.x != 0) return;
  int neg = 0, nul = 0, pos = 0;
  double t = 0.0;
  for(int k = 0; k < N; k++) {
    double d = dg[k];
    if(mode == HB_FACT_BUNCH_KAUFMAN && ipiv[k] <= 0) {
      if(t == 0.0) {
        if(k + 1 < N) {
          t = fabs(sub[k]);
          d = (d / t) * dg[k + 1] - t;
        }
      } else {
        d = t;
        t = 0.0;
      }
    }
    if(d < -1e-14) neg++;
    else if(d < 1e-14) nul++;
    else pos++;
  }
  out[0] = neg; out[1] = nul; out[2] = pos;
}

// LDL^T (no pivoting) single-rhs solve: forward (unit L), D, backward.
__global__ void __launch_bounds__(SOLVE_THREADS)
k_ldl_solve(const double* __restrict__ F, int ldf, int N, double* __restrict__ x)
{
  __shared__ double sm[32];
  __shared__ double sd[32][33];
  dev_forward(F, ldf, N, x, true, sd);
  for(int i = threadIdx.x; i < N; i += SOLVE_THREADS) x[i] /= LC(F, ldf, i, i);
  __syncthreads();
  dev_backward(F, ldf, N, x, true, sm, sd);
}
__global__ void __launch_bounds__(SOLVE_THREADS)
k_chol_solve(const double* __restrict__ F, int ldf, int N, double* __restrict__ x)
{
  __shared__ double sm[32];
  __shared__ double sd[32][33];
  dev_forward(F, ldf, N, x, false, sd);
  dev_backward(F, ldf, N, x, false, sm, sd);
}

} // namespace

// ---------------------------------------------------------------------------------------------------------
// internal API
// ---------------------------------------------------------------------------------------------------------
constexpr size_t TRAILING_SMEM = sizeof(double) * 2 * NB * TLD;
static bool g_trailing_attr = false;

int hb_dense_chol_coop(hb_ctx* c, int N, double* A, int lda, int* info_dev, double* invd, bool* used);
int hb_dense_chol_diag_inverses(hb_ctx* c, int N, const double* F, int ldf, double* invd);
bool hb_dense_coop_available(hb_ctx* c);
int hb_dense_spd_solve_coop(hb_ctx* c, int N, const double* F, int ldf, const double* invd, const double* s, const double* Nref, int ldn,
                            const double* rhs, double* x, double* work, double tol, int max_refine, double* stats_dev, bool* used);

int hb_dense_factor_blocked(hb_ctx* c, int N, double* A, int lda, bool ldl, double* Wpanel /* NB*N doubles if ldl */, int* info_dev)
{
  if(!ldl) { // small SPD systems: one cooperative launch instead of two launches per panel (hb_chol_coop.cu)
    bool used = false;
    HB_CHECK(hb_dense_chol_coop(c, N, A, lda, info_dev, nullptr, &used));
    if(used) return HB_OK;
  }
  if(!g_trailing_attr) {
    HB_CUDA(cudaFuncSetAttribute(k_trailing, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRAILING_SMEM));
    g_trailing_attr = true;
  }
  HB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), c->stream));
  for(int k0 = 0; k0 < N; k0 += NB) {
    const int nb = N - k0 < NB ? N - k0 : NB;
    const int rest = N - k0 - nb;
    int blocks = (rest + PANEL_THREADS - 1) / PANEL_THREADS;
    if(blocks < 1) blocks = 1;
    if(ldl) k_panel<true><<<blocks, PANEL_THREADS, 0, c->stream>>>(A, lda, N, k0, nb, Wpanel, N, info_dev);
    else k_panel<false><<<blocks, PANEL_THREADS, 0, c->stream>>>(A, lda, N, k0, nb, nullptr, 0, info_dev);
    HB_LAUNCHED();
    if(rest > 0) {
      const int nt = (rest + TT - 1) / TT;
      const int ntiles = nt * (nt + 1) / 2;
      const double* Q = A + (size_t)k0 * lda; // factor panel rows: Q[p][i] = Lc(i, k0+p)
      const double* P = ldl ? Wpanel : Q;
      k_trailing<<<ntiles, 128, TRAILING_SMEM, c->stream>>>(A, lda, N, k0 + nb, P, ldl ? (long long)N : (long long)lda, Q, lda, nb);
      HB_LAUNCHED();
    }
  }
  return HB_OK;
}

int hb_dense_sytf2(hb_ctx* c, int N, double* A, int lda, int* ipiv_dev, int* info_dev)
{
  if(N == 0) return HB_OK;
  k_sytf2<<<1, BK_THREADS, 0, c->stream>>>(A, lda, N, ipiv_dev, info_dev);
  HB_LAUNCHED();
  return HB_OK;
}

int hb_dense_sytrs(hb_ctx* c, int N, const double* A, int lda, const int* ipiv_dev, double* B, int ldb, int nrhs)
{
  if(N == 0 || nrhs == 0) return HB_OK;
  if(nrhs >= 8 || N <= 64) {
    k_sytrs_per_rhs<<<(nrhs + 63) / 64, 64, 0, c->stream>>>(A, lda, N, ipiv_dev, B, ldb, nrhs);
    HB_LAUNCHED();
  } else {
    for(int r = 0; r < nrhs; r++) {
      k_sytrs_cta<<<1, SOLVE_THREADS, 0, c->stream>>>(A, lda, N, ipiv_dev, B + (size_t)r * ldb);
      HB_LAUNCHED();
    }
  }
  return HB_OK;
}

int hb_dense_inertia(hb_ctx* c, int N, const double* A, int lda, const int* ipiv_dev, int mode, int* out3_dev)
{
  HB_CHECK(hb_ws_reserve(c, sizeof(double) * 2 * (size_t)(N > 0 ? N : 1) + 256));
  k_inertia<<<1, 1024, 0, c->stream>>>(A, lda, N, ipiv_dev, mode, out3_dev, reinterpret_cast<double*>(reinterpret_cast<char*>(c->ws) + 256));
  HB_LAUNCHED();
  return HB_OK;
}

int hb_dense_tri_solve(hb_ctx* c, int N, const double* F, int ldf, bool ldl, double* x)
{
  if(N == 0) return HB_OK;
  if(ldl) k_ldl_solve<<<1, SOLVE_THREADS, 0, c->stream>>>(F, ldf, N, x);
  else k_chol_solve<<<1, SOLVE_THREADS, 0, c->stream>>>(F, ldf, N, x);
  HB_LAUNCHED();
  return HB_OK;
}

int hb_dense_equilibrate(hb_ctx* c, int N, const double* Nfull, int ldn, double* F, int ldf, double* s)
{
  if(N == 0) return HB_OK;
  long long total = (long long)N * N;
  int g = (int)((total + 255) / 256 < (long long)c->num_sms * 8 ? (total + 255) / 256 : (long long)c->num_sms * 8);
  k_equilibrate<<<g, 256, 0, c->stream>>>(Nfull, ldn, N, F, ldf, s);
  HB_LAUNCHED();
  return HB_OK;
}

// SPD factorization that also keeps the 16 x 16 diagonal inverses for the cooperative solve (invd: HB_CHOL_INV_DOUBLES(N) doubles);
// *have_inv tells whether they were produced (small / large N use the multi-launch path and the one-CTA solve)
static hb_big g_chol_big[16]; // look-ahead state (panel stream, scratch) of the large condensed systems, one per device

int hb_dense_chol_with_inverses(hb_ctx* c, int N, double* A, int lda, int* info_dev, double* invd, bool* have_inv)
{
  *have_inv = false;
  HB_CHECK(hb_dense_chol_coop(c, N, A, lda, info_dev, invd, have_inv));
  if(*have_inv) return HB_OK;
  // beyond the single-launch cooperative kernel: the look-ahead Cholesky of hb_dense_big.cu (config 4: m = 4000 per condensed system)
  if(N > 2048 && (lda & 1) == 0 && (reinterpret_cast<uintptr_t>(A) & 15u) == 0 && c->device < 16) {
    HB_CHECK(hb_big_factor(c, &g_chol_big[c->device], N, A, lda, false, info_dev));
  } else
  HB_CHECK(hb_dense_factor_blocked(c, N, A, lda, false, nullptr, info_dev));
  if(N > 64 && invd && hb_dense_coop_available(c)) { // large N: multi-launch factor, but the solve can still be cooperative
    HB_CHECK(hb_dense_chol_diag_inverses(c, N, A, lda, invd));
    *have_inv = true;
  }
  return HB_OK;
}

int hb_dense_spd_solve_refine2(hb_ctx* c, int N, const double* F, int ldf, const double* invd, const double* s, const double* Nref, int ldn,
                               const double* rhs, double* x, double* work2N2, double tol, int max_refine, double* stats_dev)
{
  if(N == 0) return HB_OK;
  if(invd) {
    bool used = false;
    HB_CHECK(hb_dense_spd_solve_coop(c, N, F, ldf, invd, s, Nref, ldn, rhs, x, work2N2, tol, max_refine, stats_dev, &used));
    if(used) return HB_OK;
  }
  return hb_dense_spd_solve_refine(c, N, F, ldf, s, Nref, ldn, rhs, x, work2N2, tol, max_refine, stats_dev);
}

int hb_dense_spd_solve_refine(hb_ctx* c, int N, const double* F, int ldf, const double* s, const double* Nref, int ldn, const double* rhs,
                              double* x, double* work2N, double tol, int max_refine, double* stats_dev)
{
  if(N == 0) return HB_OK;
  k_spd_solve_refine<<<1, SOLVE_THREADS, 0, c->stream>>>(F, ldf, N, s, Nref, ldn, rhs, x, work2N, tol, max_refine, stats_dev);
  HB_LAUNCHED();
  return HB_OK;
}
