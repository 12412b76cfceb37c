// Blocked Bunch-Kaufman LDL^T (DSYTRF 'L' structure: DLASYF panels + rank-kb trailing updates) on device.
// Role: hiopLinSolverSymDenseLapack::matrixChanged's DSYTRF (src/LinAlg/hiopLinSolverSymDenseLapack.hpp:90-102) and
// hiopLinSolverSymDenseMagmaBuKa's magma_dsytrf_gpu (src/LinAlg/hiopLinSolverSymDenseMagma.cpp:151).
//
// The panel (<= 64 columns) is factorized by ONE CTA with LAPACK's pivot rule (same pivots as DSYTF2), keeping
// W = L*D of the panel in a scratch buffer; the trailing matrix is then updated A22 -= L21 * W21^T on the DMMA pipe by
// a grid of 64x64 tiles. No host synchronisation inside the loop: the number of columns a panel managed to
// factorize (63 or 64, a 2x2 pivot may not straddle the panel edge) lives in a device-side state word that the
// next kernels read.
#include "hb_common.cuh"
#include "hb_dense.cuh"

namespace {

#define LC(A, lda, i, j) (A)[(size_t)(j) * (lda) + (i)]
#define WC(W, ldw, i, c) (W)[(size_t)(c) * (ldw) + (i)]

constexpr int NBK = 64;
constexpr int PT = 1024;
#define BK_ALPHA 0.6403882032022076

struct ArgMax
{
  double v;
  int i;
};
__device__ __forceinline__ ArgMax amax_comb(ArgMax a, ArgMax b)
{
  if(b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ ArgMax cta_argmax(ArgMax a, ArgMax* sm)
{
#pragma unroll
  for(int o = 16; o > 0; o >>= 1) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = amax_comb(a, b);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if(lane == 0) sm[warp] = a;
  __syncthreads();
  ArgMax r = sm[0];
  for(int w = 1; w < PT / 32; w++) r = amax_comb(r, sm[w]);
  return r;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// state[0] = k0 of the current panel, state[1] = kb factorized by the last panel, state[2] = info (first zero pivot, 1-based)
__global__ void __launch_bounds__(PT)
k_lasyf_panel(double* __restrict__ A, int lda, int N, double* __restrict__ W, int ldw, int* __restrict__ ipiv, int* __restrict__ state)
{
  __shared__ ArgMax sm[32];
  __shared__ double wrow[NBK];
  const int tid = threadIdx.x;
  const int big = 0x7fffffff;
  const int k0 = state[0];
  if(k0 >= N) {
    if(tid == 0) state[1] = 0;
    return;
  }
  const int ns = N - k0;
  const bool last = ns <= NBK;
  int linfo = 0;
  int k = k0;
  while(true) {
    const int kl = k - k0;
    if(k >= N) break;
    if(!last && kl >= NBK - 1) break;
    // --- W(k:N,kl) = A(k:N,k) - A(k:N,k0:k-1) * W(k,0:kl-1)^T
    __syncthreads();
    for(int c = tid; c < kl; c += PT) wrow[c] = WC(W, ldw, k, c);
    __syncthreads();
    for(int i = k + tid; i < N; i += PT) {
      double v = LC(A, lda, i, k), v1 = 0.0, v2 = 0.0, v3 = 0.0;
      int c = 0;
#pragma unroll 2
      for(; c + 4 <= kl; c += 4) { // four independent accumulators: the loads of a batch are in flight together
        v -= LC(A, lda, i, k0 + c) * wrow[c];
        v1 -= LC(A, lda, i, k0 + c + 1) * wrow[c + 1];
        v2 -= LC(A, lda, i, k0 + c + 2) * wrow[c + 2];
        v3 -= LC(A, lda, i, k0 + c + 3) * wrow[c + 3];
      }
      for(; c < kl; c++) v -= LC(A, lda, i, k0 + c) * wrow[c];
      WC(W, ldw, i, kl) = (v + v1) + (v2 + v3);
    }
    __syncthreads();
    int kstep = 1, kp = k;
    const double absakk = fabs(WC(W, ldw, k, kl));
    int imax = k;
    double colmax = 0.0;
    if(k < N - 1) {
      ArgMax a{-1.0, big};
      for(int i = k + 1 + tid; i < N; i += PT) a = amax_comb(a, ArgMax{fabs(WC(W, ldw, i, kl)), i});
      a = cta_argmax(a, sm);
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
        // column imax (updated) into W(:,kl+1)
        __syncthreads();
        for(int c = tid; c < kl; c += PT) wrow[c] = WC(W, ldw, imax, c);
        __syncthreads();
        for(int i = k + tid; i < N; i += PT) {
          double v = i < imax ? LC(A, lda, imax, i) : LC(A, lda, i, imax), v1 = 0.0, v2 = 0.0, v3 = 0.0;
          int c = 0;
#pragma unroll 2
          for(; c + 4 <= kl; c += 4) {
            v -= LC(A, lda, i, k0 + c) * wrow[c];
            v1 -= LC(A, lda, i, k0 + c + 1) * wrow[c + 1];
            v2 -= LC(A, lda, i, k0 + c + 2) * wrow[c + 2];
            v3 -= LC(A, lda, i, k0 + c + 3) * wrow[c + 3];
          }
          for(; c < kl; c++) v -= LC(A, lda, i, k0 + c) * wrow[c];
          WC(W, ldw, i, kl + 1) = (v + v1) + (v2 + v3);
        }
        __syncthreads();
        ArgMax a{-1.0, big};
        for(int i = k + tid; i < N; i += PT)
          if(i != imax) a = amax_comb(a, ArgMax{fabs(WC(W, ldw, i, kl + 1)), i});
        a = cta_argmax(a, sm);
        const double rowmax = a.v;
        if(absakk >= BK_ALPHA * colmax * (colmax / rowmax)) {
          kp = k;
        } else if(fabs(WC(W, ldw, imax, kl + 1)) >= BK_ALPHA * rowmax) {
          kp = imax;
          __syncthreads();
          for(int i = k + tid; i < N; i += PT) WC(W, ldw, i, kl) = WC(W, ldw, i, kl + 1);
          __syncthreads();
        } else {
          kp = imax;
          kstep = 2;
        }
      }
      const int kk = k + kstep - 1, kkl = kk - k0;
      __syncthreads();
      if(kp != kk) {
        // copy the non-updated column kk into position kp of the trailing submatrix
        if(tid == 0) LC(A, lda, kp, kp) = LC(A, lda, kk, kk);
        for(int i = kk + 1 + tid; i < kp; i += PT) LC(A, lda, kp, i) = LC(A, lda, i, kk);
        for(int i = kp + 1 + tid; i < N; i += PT) LC(A, lda, i, kp) = LC(A, lda, i, kk);
        // swap rows kk and kp in the panel's finished columns of A and in W(.,0:kkl)
        for(int c = tid; c < kl; c += PT) {
          const double t = LC(A, lda, kk, k0 + c);
          LC(A, lda, kk, k0 + c) = LC(A, lda, kp, k0 + c);
          LC(A, lda, kp, k0 + c) = t;
        }
        for(int c = tid; c <= kkl; c += PT) {
          const double t = WC(W, ldw, kk, c);
          WC(W, ldw, kk, c) = WC(W, ldw, kp, c);
          WC(W, ldw, kp, c) = t;
        }
        __syncthreads();
      }
      if(kstep == 1) {
        const double akk = WC(W, ldw, k, kl);
        const double r1 = 1.0 / akk;
        for(int i = k + tid; i < N; i += PT) {
          const double w = WC(W, ldw, i, kl);
          LC(A, lda, i, k) = (i == k) ? w : w * r1;
        }
      } else {
        if(k < N - 2) {
          double d21 = WC(W, ldw, k + 1, kl);
          const double d11 = WC(W, ldw, k + 1, kl + 1) / d21;
          const double d22 = WC(W, ldw, k, kl) / d21;
          const double t = 1.0 / (d11 * d22 - 1.0);
          d21 = t / d21;
          for(int j = k + 2 + tid; j < N; j += PT) {
            const double wj0 = WC(W, ldw, j, kl), wj1 = WC(W, ldw, j, kl + 1);
            LC(A, lda, j, k) = d21 * (d11 * wj0 - wj1);
            LC(A, lda, j, k + 1) = d21 * (d22 * wj1 - wj0);
          }
        }
        if(tid == 0) {
          LC(A, lda, k, k) = WC(W, ldw, k, kl);
          LC(A, lda, k + 1, k) = WC(W, ldw, k + 1, kl);
          LC(A, lda, k + 1, k + 1) = WC(W, ldw, k + 1, kl + 1);
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
  if(tid == 0) {
    state[1] = k - k0;
    if(linfo != 0 && state[2] == 0) state[2] = linfo;
  }
}

// A22 -= L21 * W21^T on the lower triangle, r0 = k0 + kb (read from the device state). 64x64 tiles.
constexpr int TT = 64;
constexpr int TLD = TT + 4;
__global__ void __launch_bounds__(128)
k_bk_trailing(double* __restrict__ A, int lda, int N, const double* __restrict__ W, int ldw, const int* __restrict__ state)
{
  extern __shared__ __align__(16) unsigned char tsm[];
  double (*sP)[TLD] = reinterpret_cast<double (*)[TLD]>(tsm);
  double (*sQ)[TLD] = sP + NBK;
  const int k0 = state[0], kb = state[1];
  const int r0 = k0 + kb;
  if(kb == 0 || r0 >= N) return;
  const int nt = (N - r0 + TT - 1) / TT;
  int t = blockIdx.x, ti = 0;
  while(t >= ti + 1) { t -= ti + 1; ti++; }
  const int tj = t;
  if(ti >= nt) return;
  const int i0 = r0 + ti * TT, j0 = r0 + tj * TT;
  const int tid = threadIdx.x;
  const int kpad = ((kb + 3) / 4) * 4;
  for(int e = tid; e < kpad * TT; e += 128) {
    const int p = e / TT, c = e % TT;
    const bool vp = (p < kb) && (i0 + c < N), vq = (p < kb) && (j0 + c < N);
    const unsigned sp = (unsigned)__cvta_generic_to_shared(&sP[p][c]), sq = (unsigned)__cvta_generic_to_shared(&sQ[p][c]);
    const double* gp = vp ? &LC(A, lda, i0 + c, k0 + p) : A; // L21
    const double* gq = vq ? &WC(W, ldw, j0 + c, p) : W;      // W21
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sp), "l"(gp), "r"(vp ? 8 : 0));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sq), "l"(gq), "r"(vq ? 8 : 0));
  }
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int wi = warp & 1, wj = warp >> 1;
  const int g = lane >> 2, t4 = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for(int a = 0; a < 4; a++)
#pragma unroll
    for(int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
  for(int kk = 0; kk < kpad / 4; kk++) {
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

// Puts L21 of the finished panel in LAPACK's standard form (partial undo of the row interchanges, DLASYF label 120)
// and advances the state to the next panel.
__global__ void k_lasyf_finish(double* __restrict__ A, int lda, int N, const int* __restrict__ ipiv, int* __restrict__ state)
{
  const int k0 = state[0], kb = state[1];
  if(kb == 0) return;
  const int kend = k0 + kb;
  int j = kend - 1;
  while(j >= k0) {
    const int jj = j;
    int jp = ipiv[j];
    if(jp < 0) { jp = -jp; j -= 1; }
    jp -= 1;
    j -= 1;
    const int ncols = j - k0 + 1;
    if(jp != jj && ncols >= 1) {
      for(int c = threadIdx.x; c < ncols; c += blockDim.x) {
        const double t = LC(A, lda, jp, k0 + c);
        LC(A, lda, jp, k0 + c) = LC(A, lda, jj, k0 + c);
        LC(A, lda, jj, k0 + c) = t;
      }
    }
    __syncthreads();
    if(j <= k0) break;
  }
  __syncthreads();
  if(threadIdx.x == 0) state[0] = kend;
}

bool g_bk_attr = false;

} // namespace

int hb_dense_sytrf_blocked(hb_ctx* c, int N, double* A, int lda, int* ipiv_dev, double* Wpanel, int* info_dev)
{
  if(N == 0) return HB_OK;
  const size_t smem = sizeof(double) * 2 * NBK * TLD;
  if(!g_bk_attr) {
    HB_CUDA(cudaFuncSetAttribute(k_bk_trailing, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    g_bk_attr = true;
  }
  // device state lives at the tail of the ipiv buffer's companion: reuse info_dev[0..2]? info_dev has 4 ints: [0]=info,[1..3]=inertia.
  // Use a small dedicated allocation from the workspace instead.
  HB_CHECK(hb_ws_reserve(c, 64));
  int* state = reinterpret_cast<int*>(c->ws);
  HB_CUDA(cudaMemsetAsync(state, 0, sizeof(int) * 4, c->stream));
  const int max_panels = (N + (NBK - 1) - 1) / (NBK - 1) + 1;
  for(int p = 0; p < max_panels; p++) {
    const int k0_min = p * (NBK - 1); // a panel advances by at least NBK-1 columns
    if(k0_min >= N) break;
    k_lasyf_panel<<<1, PT, 0, c->stream>>>(A, lda, N, Wpanel, N, ipiv_dev, state);
    HB_LAUNCHED();
    const int rest_max = N - k0_min - (NBK - 1);
    if(rest_max > 0) {
      const int nt = (rest_max + TT - 1) / TT;
      k_bk_trailing<<<nt * (nt + 1) / 2, 128, smem, c->stream>>>(A, lda, N, Wpanel, N, state);
      HB_LAUNCHED();
    }
    k_lasyf_finish<<<1, 64, 0, c->stream>>>(A, lda, N, ipiv_dev, state);
    HB_LAUNCHED();
  }
  HB_CUDA(cudaMemcpyAsync(info_dev, state + 2, sizeof(int), cudaMemcpyDeviceToDevice, c->stream));
  return HB_OK;
}
