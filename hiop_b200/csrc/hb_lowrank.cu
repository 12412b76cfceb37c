// hiopKKTLinSysLowRank + hiopHessianLowRank on device (the quasi-Newton KKT path).
// Reference: src/Optimization/hiopKKTLinSys.cpp:1031-1350, src/Optimization/hiopHessianLowRank.cpp:221-630, 974-1059,
// hiopKKTLinSysCompressedXYcYd::computeDirections hiopKKTLinSys.cpp:585-691, compute_directions_for_full_space :218-309.
#include "hb_common.cuh"
#include "hb_dense.cuh"
#include "hb_lowrank.cuh"
#include <cstdlib>

int hb_syrk_rows(hb_ctx* c, int M, long long K, const double* const* rowptr_dev, bool aligned16, const double* d, double* C, int ldc);
int hb_syrk_rows_ozaki(hb_ctx* c, int M, long long K, const double* const* rowptr_dev, bool rows_aligned16, const double* d, double* C, int ldc, int S,
                       const double* dot_x, double* dot_out);

namespace {

constexpr int ET = 256; // elementwise / streaming kernels

inline int stream_grid(hb_ctx* c, long long items)
{
  long long g = (items + ET - 1) / ET;
  long long cap = (long long)c->num_sms * 8;
  if(g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

// ---- a1: Dx, DhInv in one pass (hiopKKTLinSys.cpp:1073-1078 + hiopHessianLowRank.cpp:223-229; 7 passes there) ----------
// Same operation order as the reference so the result is bit-identical: Dx = 0 + zl/sxl (+ zu/sxu); DhInv = 1/(sigma + Dx).
__global__ void __launch_bounds__(ET)
k_update_x(long long n, const double* __restrict__ zl, const double* __restrict__ sxl, const double* __restrict__ zu,
           const double* __restrict__ sxu, const double* __restrict__ ixl, const double* __restrict__ ixu, double sigma,
           double* __restrict__ Dx, double* __restrict__ DhInv)
{
  const long long stride = (long long)gridDim.x * ET;
  for(long long i = (long long)blockIdx.x * ET + threadIdx.x; i < n; i += stride) {
    double d = 0.0;
    if(ixl[i] == 1.0) d = __dadd_rn(d, __ddiv_rn(zl[i], sxl[i]));
    if(ixu[i] == 1.0) d = __dadd_rn(d, __ddiv_rn(zu[i], sxu[i]));
    Dx[i] = d;
    if(DhInv) DhInv[i] = __ddiv_rn(1.0, __dadd_rn(sigma, d));
  }
}
// d-side: Dd = vl/sdl|idl + vu/sdu|idu, Dd_inv = 1/Dd (hiopKKTLinSys.cpp:1081-1088)
__global__ void k_update_d(int mi, const double* __restrict__ vl, const double* __restrict__ sdl, const double* __restrict__ vu,
                           const double* __restrict__ sdu, const double* __restrict__ idl, const double* __restrict__ idu,
                           double* __restrict__ Dd, double* __restrict__ Dd_inv)
{
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < mi; i += gridDim.x * blockDim.x) {
    double d = 0.0;
    if(idl[i] == 1.0) d = __dadd_rn(d, __ddiv_rn(vl[i], sdl[i]));
    if(idu[i] == 1.0) d = __dadd_rn(d, __ddiv_rn(vu[i], sdu[i]));
    Dd[i] = d;
    Dd_inv[i] = __ddiv_rn(1.0, d);
  }
}

// ---- a8: V from the blocks of C_aug = [J;S;Y] DhInv [J;S;Y]^T (updateInternalBFGSRepresentation, :400-485) ----------------
__global__ void k_build_V(int m, int l, double sigma, const double* __restrict__ C, int ldc, const double* __restrict__ SSt,
                          const double* __restrict__ L, const double* __restrict__ D, double* __restrict__ V)
{
  const int n2 = 2 * l;
  for(int e = blockIdx.x * blockDim.x + threadIdx.x; e < n2 * n2; e += gridDim.x * blockDim.x) {
    const int a = e / n2, b = e % n2;
    double v;
    if(a < l && b < l) v = sigma * sigma * C[(size_t)(m + a) * ldc + m + b] - sigma * SSt[a * l + b];
    else if(a < l && b >= l) v = sigma * C[(size_t)(m + a) * ldc + m + b] - L[a * l + (b - l)];
    else if(a >= l && b < l) v = sigma * C[(size_t)(m + b) * ldc + m + a] - L[b * l + (a - l)];
    else v = C[(size_t)(m + a) * ldc + m + b] + (a == b ? D[a - l] : 0.0);
    V[a * n2 + b] = v;
  }
}
// M = [[sigma S^T S, L],[L^T, -D]] of the compact (direct) BFGS representation, used by hess_times_vec
__global__ void k_build_Mdirect(int l, double sigma, const double* __restrict__ SSt, const double* __restrict__ L,
                                const double* __restrict__ D, double* __restrict__ Mm)
{
  const int n2 = 2 * l;
  for(int e = blockIdx.x * blockDim.x + threadIdx.x; e < n2 * n2; e += gridDim.x * blockDim.x) {
    const int a = e / n2, b = e % n2;
    double v;
    if(a < l && b < l) v = sigma * SSt[a * l + b];
    else if(a < l && b >= l) v = L[a * l + (b - l)];
    else if(a >= l && b < l) v = L[b * l + (a - l)];
    else v = (a == b ? -D[a - l] : 0.0);
    Mm[a * n2 + b] = v;
  }
}
// U = [S1 Y1] (m x 2l): S1 = sigma * C[0:m, m:m+l], Y1 = C[0:m, m+l:m+2l]; Z = copy of U (rhs of the V solve)
__global__ void k_build_U(int m, int l, double sigma, const double* __restrict__ C, int ldc, double* __restrict__ U, double* __restrict__ Z)
{
  const int n2 = 2 * l;
  for(int e = blockIdx.x * blockDim.x + threadIdx.x; e < m * n2; e += gridDim.x * blockDim.x) {
    const int i = e / n2, q = e % n2;
    double v = C[(size_t)i * ldc + m + q];
    if(q < l) v *= sigma;
    U[e] = v;
    Z[e] = v;
  }
}
// N = W0 - U Z^T + blkdiag(0, Dd_inv) (hiopHessianLowRank.cpp:608-618 + hiopKKTLinSys.cpp:1135); computed for i <= j, mirrored.
__global__ void k_form_N(int m, int meq, int l, const double* __restrict__ C, int ldc, const double* __restrict__ U, const double* __restrict__ Z,
                         const double* __restrict__ Dd_inv, double* __restrict__ Nm)
{
  const int n2 = 2 * l;
  for(long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)m * m; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / m), j = (int)(e % m);
    if(j < i) continue;
    double v = C[(size_t)i * ldc + j];
    double corr = 0.0;
    for(int q = 0; q < n2; q++) corr += U[(size_t)i * n2 + q] * Z[(size_t)j * n2 + q];
    v -= corr;
    if(i == j && i >= meq) v += Dd_inv[i - meq];
    Nm[(size_t)i * m + j] = v;
    Nm[(size_t)j * m + i] = v;
  }
}

// ---- all-reduce of the symmetric C_aug as its packed upper triangle (hiopHessianLowRank.cpp:590-591 sends the full m x m buffer) ----
__global__ void k_pack_upper(int M, const double* __restrict__ C, int ldc, double* __restrict__ tri)
{
  const long long tot = (long long)M * (M + 1) / 2;
  for(long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    // row i of the upper triangle starts at i*M - i(i-1)/2
    int i = (int)((2.0 * M + 1.0 - sqrt((2.0 * M + 1.0) * (2.0 * M + 1.0) - 8.0 * (double)e)) * 0.5);
    while((long long)i * M - (long long)i * (i - 1) / 2 > e) i--;
    while((long long)(i + 1) * M - (long long)(i + 1) * i / 2 <= e) i++;
    const int j = i + (int)(e - ((long long)i * M - (long long)i * (i - 1) / 2));
    tri[e] = C[(size_t)i * ldc + j];
  }
}
__global__ void k_unpack_upper(int M, double* __restrict__ C, int ldc, const double* __restrict__ tri)
{
  const long long tot = (long long)M * (M + 1) / 2;
  for(long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    int i = (int)((2.0 * M + 1.0 - sqrt((2.0 * M + 1.0) * (2.0 * M + 1.0) - 8.0 * (double)e)) * 0.5);
    while((long long)i * M - (long long)i * (i - 1) / 2 > e) i--;
    while((long long)(i + 1) * M - (long long)(i + 1) * i / 2 <= e) i++;
    const int j = i + (int)(e - ((long long)i * M - (long long)i * (i - 1) / 2));
    const double v = tri[e];
    C[(size_t)i * ldc + j] = v;
    C[(size_t)j * ldc + i] = v;
  }
}

// ---- multi-dot: out[q] = sum_k R_q[k] * w[k] * x[k] * scale_q, q < nq rows (two-stage, deterministic) ------------------------
// rows q < l come from S (scale sigma_s), rows q >= l from Y (scale 1). w may be null (=1).
constexpr int MD_CH = 8;
__global__ void __launch_bounds__(ET)
k_multidot_partial(long long n, int l, const double* __restrict__ S, const double* __restrict__ Y, long long ld, const double* __restrict__ w,
                   const double* __restrict__ x, int q0, double* __restrict__ partial /* [grid][2l] */)
{
  __shared__ double sm[ET / 32];
  const int n2 = 2 * l;
  double acc[MD_CH];
#pragma unroll
  for(int c = 0; c < MD_CH; c++) acc[c] = 0.0;
  const long long stride = (long long)gridDim.x * ET;
  for(long long k = (long long)blockIdx.x * ET + threadIdx.x; k < n; k += stride) {
    const double t = w ? w[k] * x[k] : x[k];
#pragma unroll
    for(int c = 0; c < MD_CH; c++) {
      const int q = q0 + c;
      if(q < n2) {
        const double* row = q < l ? S + (size_t)q * ld : Y + (size_t)(q - l) * ld;
        acc[c] += row[k] * t;
      }
    }
  }
#pragma unroll
  for(int c = 0; c < MD_CH; c++) {
    const double r = hb_block_sum<ET>(acc[c], sm);
    if(threadIdx.x == 0 && q0 + c < n2) partial[(size_t)blockIdx.x * n2 + q0 + c] = r;
  }
}
// one CTA per output: 128 threads stride over the per-CTA partials, fixed-order block reduction (a single thread walking
// all ~1200 partials took 98 us)
__global__ void __launch_bounds__(128)
k_multidot_final(int np, int l, double sigma_s, const double* __restrict__ partial, double* __restrict__ out)
{
  __shared__ double sm[4];
  const int n2 = 2 * l;
  const int q = blockIdx.x;
  double s = 0.0;
  for(int p = threadIdx.x; p < np; p += 128) s += partial[(size_t)p * n2 + q];
  s = hb_block_sum<128>(s, sm);
  if(threadIdx.x == 0) out[q] = q < l ? s * sigma_s : s;
}
// x[k] = w[k] * (r[k] - sigma*sum_q S_q[k] p_q - sum_q Y_q[k] p_{l+q})   (hiopHessianLowRank::solve steps 4-5, :526-535)
// general form: out = beta*out + alpha*( base[k]*r[k] ... ) handled by flags below
__global__ void __launch_bounds__(ET)
k_lowrank_apply(long long n, int l, double sigma, const double* __restrict__ S, const double* __restrict__ Y, long long ld,
                const double* __restrict__ p /* 2l */, const double* __restrict__ w /* n or null */, const double* __restrict__ r,
                double diag_scale, const double* __restrict__ diag_add /* n or null */, double beta, double alpha, double* __restrict__ out)
{
  // t = (diag_scale + diag_add[k]) * r[k]  (when w == null)   or   t = r[k] (when w != null)
  // v = t - (sigma*S^T p_s + Y^T p_y)[k];   if w: v *= w[k];   out = beta*out + alpha*v
  extern __shared__ double sp[];
  for(int q = threadIdx.x; q < 2 * l; q += ET) sp[q] = p[q];
  __syncthreads();
  const long long stride = (long long)gridDim.x * ET;
  for(long long k = (long long)blockIdx.x * ET + threadIdx.x; k < n; k += stride) {
    double ss = 0.0, sy = 0.0;
    for(int q = 0; q < l; q++) {
      ss += S[(size_t)q * ld + k] * sp[q];
      sy += Y[(size_t)q * ld + k] * sp[l + q];
    }
    const double corr = sigma * ss + sy;
    double v;
    if(w) v = w[k] * (r[k] - corr);
    else v = (diag_scale + (diag_add ? diag_add[k] : 0.0)) * r[k] - corr;
    out[k] = (beta == 0.0 ? 0.0 : beta * out[k]) + alpha * v;
  }
}

// ---- a13: J*x (rows) and J^T*y (columns) -----------------------------------------------------------------------------------
constexpr int GR_THREADS = 512;
constexpr int GR_CHUNK = 2048; // columns per CTA
__global__ void __launch_bounds__(GR_THREADS)
k_gemv_rows_partial(int m, long long n, const double* __restrict__ A, long long lda, const double* __restrict__ x,
                    double* __restrict__ partial /* [nchunks][m] */)
{
  __shared__ double sx[GR_CHUNK];
  const long long k0 = (long long)blockIdx.x * GR_CHUNK;
  const int len = (int)min((long long)GR_CHUNK, n - k0);
  for(int k = threadIdx.x; k < GR_CHUNK; k += GR_THREADS) sx[k] = k < len ? x[k0 + k] : 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0) && ((k0 & 1) == 0);
  // blockIdx.y interleaves the rows among gridDim.y CTAs of the same column chunk (short local column ranges)
  for(int i = warp + (GR_THREADS / 32) * blockIdx.y; i < m; i += (GR_THREADS / 32) * gridDim.y) {
    const double* row = A + (size_t)i * lda + k0;
    double acc = 0.0;
    if(vec) {
      const double2* r2 = reinterpret_cast<const double2*>(row);
      const int len2 = len >> 1;
#pragma unroll 8
      for(int k = lane; k < len2; k += 32) {
        const double2 v = r2[k];
        acc += v.x * sx[2 * k] + v.y * sx[2 * k + 1];
      }
      if((len & 1) && lane == 0) acc += row[len - 1] * sx[len - 1];
    } else {
#pragma unroll 8
      for(int k = lane; k < len; k += 32) acc += row[k] * sx[k];
    }
    acc = hb_warp_sum(acc);
    if(lane == 0) partial[(size_t)blockIdx.x * m + i] = acc;
  }
}
// y[i] = beta*y[i] + alpha*sum_chunks partial[c][i]: 32 rows x 8 chunk-classes per CTA (coalesced along i), fixed-order combine
__global__ void __launch_bounds__(256)
k_gemv_rows_final(int m, int nchunks, const double* __restrict__ partial, double beta, double* __restrict__ y, double alpha)
{
  __shared__ double sm[8][33];
  const int ri = threadIdx.x & 31, part = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ri;
  double s = 0.0;
  if(i < m)
    for(int c = part; c < nchunks; c += 8) s += partial[(size_t)c * m + i];
  sm[part][ri] = s;
  __syncthreads();
  if(part == 0 && i < m) {
    double t = 0.0;
#pragma unroll
    for(int p = 0; p < 8; p++) t += sm[p][ri];
    y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + alpha * t;
  }
}

// y[k] = beta*y[k] + alpha*sum_i A[i][k]*x[i]; each thread owns two adjacent columns. With RG > 1 the CTA covers ET/RG column
// pairs and its RG thread groups take interleaved rows (combined in a fixed order through shared memory): short local column
// ranges (a rank's shard of n) still fill the machine.
constexpr int GC_ROWS = 1024; // rows of x staged per pass
template <int RG>
__global__ void __launch_bounds__(ET)
k_gemv_cols(int m, long long n, const double* __restrict__ A, long long lda, const double* __restrict__ x, double beta,
            double* __restrict__ y, double alpha)
{
  constexpr int CP = ET / RG; // column pairs per CTA
  __shared__ double sx[GC_ROWS];
  __shared__ double2 red[RG > 1 ? ET : 1];
  const bool vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
  const int cp = threadIdx.x % CP, rg = threadIdx.x / CP;
  const long long k = ((long long)blockIdx.x * CP + cp) * 2;
  double a0 = 0.0, a1 = 0.0;
  for(int i0 = 0; i0 < m; i0 += GC_ROWS) {
    const int nr = min(GC_ROWS, m - i0);
    __syncthreads();
    for(int i = threadIdx.x; i < nr; i += ET) sx[i] = x[i0 + i];
    __syncthreads();
    if(k < n) {
      const double* col = A + (size_t)i0 * lda + k;
      if(vec && k + 1 < n) {
#pragma unroll 8
        for(int i = rg; i < nr; i += RG) {
          const double2 v = *reinterpret_cast<const double2*>(col + (size_t)i * lda);
          a0 += v.x * sx[i];
          a1 += v.y * sx[i];
        }
      } else {
#pragma unroll 4
        for(int i = rg; i < nr; i += RG) {
          a0 += col[(size_t)i * lda] * sx[i];
          if(k + 1 < n) a1 += col[(size_t)i * lda + 1] * sx[i];
        }
      }
    }
  }
  if(RG > 1) {
    red[threadIdx.x] = make_double2(a0, a1);
    __syncthreads();
    if(rg != 0) return;
    a0 = 0.0, a1 = 0.0;
#pragma unroll
    for(int g = 0; g < RG; g++) {
      a0 += red[g * CP + cp].x;
      a1 += red[g * CP + cp].y;
    }
  }
  if(k < n) {
    y[k] = (beta == 0.0 ? 0.0 : beta * y[k]) + alpha * a0;
    if(k + 1 < n) y[k + 1] = (beta == 0.0 ? 0.0 : beta * y[k + 1]) + alpha * a1;
  }
}

// ---- a5: rhs reduction + back-substitution (computeDirections / compute_directions_for_full_space) ---------------------------
// Operation order follows the reference exactly (intrinsics forbid FMA contraction) so that rx_tilde etc. are bit-identical.
// out = r0 + (pl? (rsl - dl*rl)/sl) - (pu? (rsu - du*ru)/su)
__global__ void __launch_bounds__(ET)
k_reduce_rhs(long long n, const double* __restrict__ r0, const double* __restrict__ rsl, const double* __restrict__ dl, const double* __restrict__ rl,
             const double* __restrict__ sl, const double* __restrict__ pl, const double* __restrict__ rsu, const double* __restrict__ du,
             const double* __restrict__ ru, const double* __restrict__ su, const double* __restrict__ pu, double* __restrict__ out)
{
  const long long stride = (long long)gridDim.x * ET;
  for(long long i = (long long)blockIdx.x * ET + threadIdx.x; i < n; i += stride) {
    double v = r0[i];
    if(pl[i] == 1.0) v = __dadd_rn(v, __ddiv_rn(__dsub_rn(rsl[i], __dmul_rn(dl[i], rl[i])), sl[i]));
    if(pu[i] == 1.0) v = __dsub_rn(v, __ddiv_rn(__dsub_rn(rsu[i], __dmul_rn(du[i], ru[i])), su[i]));
    out[i] = v;
  }
}
// ds_l = pl ? r_l + dvar : 0 ; dz_l = pl ? (rs_l - dual_l*ds_l)/s_l : 0 ; ds_u = pu ? r_u - dvar : 0 ; dz_u = pu ? (rs_u - dual_u*ds_u)/s_u : 0
__global__ void __launch_bounds__(ET)
k_recover_slack_duals(long long n, const double* __restrict__ dvar, const double* __restrict__ rl, const double* __restrict__ rsl,
                      const double* __restrict__ dual_l, const double* __restrict__ sl, const double* __restrict__ pl,
                      const double* __restrict__ ru, const double* __restrict__ rsu, const double* __restrict__ dual_u,
                      const double* __restrict__ su, const double* __restrict__ pu, double* __restrict__ dsl, double* __restrict__ dzl,
                      double* __restrict__ dsu, double* __restrict__ dzu)
{
  const long long stride = (long long)gridDim.x * ET;
  for(long long i = (long long)blockIdx.x * ET + threadIdx.x; i < n; i += stride) {
    const double dv = dvar[i];
    double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
    if(pl[i] != 0.0) {
      a = __dadd_rn(rl[i], dv);
      b = __ddiv_rn(__dsub_rn(rsl[i], __dmul_rn(dual_l[i], a)), sl[i]);
    }
    if(pu[i] != 0.0) {
      c = __dsub_rn(ru[i], dv);
      d = __ddiv_rn(__dsub_rn(rsu[i], __dmul_rn(dual_u[i], c)), su[i]);
    }
    dsl[i] = a; dzl[i] = b; dsu[i] = c; dzu[i] = d;
  }
}
// ryd_tilde = ryd + ryd2*Dd_inv
__global__ void k_axzpy_small(int n, double* __restrict__ out, const double* __restrict__ y, const double* __restrict__ x, const double* __restrict__ z)
{
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = __dadd_rn(y[i], __dmul_rn(x[i], z[i]));
}
// dd = (ryd2 + dyd)*Dd_inv
__global__ void k_recover_dd(int n, double* __restrict__ dd, const double* __restrict__ ryd2, const double* __restrict__ dyd, const double* __restrict__ Ddinv)
{
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dd[i] = __dmul_rn(__dadd_rn(ryd2[i], dyd[i]), Ddinv[i]);
}
// rhs[i] = t[i] - r[i] for the stacked [ryc; ryd]
__global__ void k_sub_stacked(int meq, int mineq, double* __restrict__ rhs, const double* __restrict__ ryc, const double* __restrict__ ryd)
{
  const int m = meq + mineq;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) rhs[i] -= (i < meq ? ryc[i] : ryd[i - meq]);
}

// rhs[i] = tdot[i] - sum_q Z[i][q] p[q] - ry[i],  p = [sigma*tdot[m..m+l); tdot[m+l..m+2l)]   (fused steps 1-2 of solveCompressed)
__global__ void k_fused_rhs(int m, int meq, int l, double sigma, const double* __restrict__ tdot, const double* __restrict__ Z,
                            const double* __restrict__ ryc, const double* __restrict__ ryd, double* __restrict__ rhs)
{
  const int n2 = 2 * l;
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    double corr = 0.0;
    for(int q = 0; q < n2; q++) corr += Z[(size_t)i * n2 + q] * (q < l ? sigma * tdot[m + q] : tdot[m + q]);
    rhs[i] = (tdot[i] - corr) - (i < meq ? ryc[i] : ryd[i - meq]);
  }
}

} // namespace

// =============================================================================================================
namespace {

int dmalloc(double** p, size_t count)
{
  if(cudaMalloc(p, sizeof(double) * (count ? count : 1)) != cudaSuccess) {
    cudaGetLastError();
    snprintf(g_hb_err, sizeof(g_hb_err), "device allocation of %zu doubles failed", count);
    return HB_ERR_ALLOC;
  }
  return HB_OK;
}

int refresh_rowptr(hb_lowrank* k)
{
  if(!k->rowptr_dirty) return HB_OK;
  hb_ctx* c = k->ctx;
  const int Ma = k->m + 2 * k->l;
  bool al = true;
  for(int i = 0; i < k->m; i++) k->rowptr_host[i] = k->J + (size_t)i * k->n;
  for(int q = 0; q < k->l; q++) {
    k->rowptr_host[k->m + q] = k->St + (size_t)q * k->n;
    k->rowptr_host[k->m + k->l + q] = k->Yt + (size_t)q * k->n;
  }
  for(int i = 0; i < Ma; i++) al = al && ((reinterpret_cast<uintptr_t>(k->rowptr_host[i]) & 15u) == 0);
  k->rows_aligned = al;
  HB_CUDA(cudaMemcpyAsync(k->rowptr_dev, k->rowptr_host, sizeof(double*) * Ma, cudaMemcpyHostToDevice, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream)); // rowptr_host may be rewritten by the next set_* call
  k->rowptr_dirty = false;
  return HB_OK;
}

// p = V^{-1} [sigma*S (w.x); Y (w.x)] style multi-dot into k->p2l (device), all-reduced
int multidot(hb_lowrank* k, const double* w, const double* x, double sigma_s)
{
  hb_ctx* c = k->ctx;
  const int n2 = 2 * k->l;
  if(n2 == 0) return HB_OK;
  const int g = k->md_grid;
  for(int q0 = 0; q0 < n2; q0 += MD_CH) {
    k_multidot_partial<<<g, ET, 0, c->stream>>>(k->n, k->l, k->St, k->Yt, k->n, w, x, q0, k->md_partial);
    HB_LAUNCHED();
  }
  k_multidot_final<<<n2, 128, 0, c->stream>>>(g, k->l, sigma_s, k->md_partial, k->p2l);
  HB_LAUNCHED();
  HB_CHECK(hb_allreduce_sum(c, k->p2l, n2));
  return HB_OK;
}

int gemv_rows(hb_lowrank* k, const double* A, int m, double beta, double* y, double alpha, const double* x)
{
  hb_ctx* c = k->ctx;
  if(m == 0) return HB_OK;
  const int nchunks = (int)((k->n + GR_CHUNK - 1) / GR_CHUNK);
  HB_CHECK(hb_ws_reserve(c, sizeof(double) * (size_t)(nchunks > 0 ? nchunks : 1) * m));
  if(nchunks > 0) {
    int rsplit = (4 * c->num_sms + nchunks - 1) / nchunks; // at least ~4 CTAs per SM in flight
    rsplit = rsplit < 1 ? 1 : (rsplit > 8 ? 8 : rsplit);
    k_gemv_rows_partial<<<dim3(nchunks, rsplit), GR_THREADS, 0, c->stream>>>(m, k->n, A, k->n, x, (double*)c->ws);
    HB_LAUNCHED();
  }
  if(c->nranks > 1) {
    // beta*y only on rank 0 before the reduction (hiopMatrixDenseRowMajor.cpp:464-467)
    k_gemv_rows_final<<<(m + 31) / 32, 256, 0, c->stream>>>(m, nchunks, (const double*)c->ws, c->rank == 0 ? beta : 0.0, y, alpha);
    HB_LAUNCHED();
    HB_CHECK(hb_allreduce_sum(c, y, m));
  } else {
    k_gemv_rows_final<<<(m + 31) / 32, 256, 0, c->stream>>>(m, nchunks, (const double*)c->ws, beta, y, alpha);
    HB_LAUNCHED();
  }
  return HB_OK;
}

int gemv_cols(hb_lowrank* k, const double* A, int m, double beta, double* y, double alpha, const double* x)
{
  hb_ctx* c = k->ctx;
  if(k->n == 0) return HB_OK;
  const long long pairs = (k->n + 1) / 2;
  const long long ctas1 = (pairs + ET - 1) / ET;
  if(ctas1 >= 8LL * c->num_sms || m < 64) k_gemv_cols<1><<<(unsigned)ctas1, ET, 0, c->stream>>>(m, k->n, A, k->n, x, beta, y, alpha);
  else if(ctas1 >= 2LL * c->num_sms) k_gemv_cols<4><<<(unsigned)((pairs + ET / 4 - 1) / (ET / 4)), ET, 0, c->stream>>>(m, k->n, A, k->n, x, beta, y, alpha);
  else k_gemv_cols<8><<<(unsigned)((pairs + ET / 8 - 1) / (ET / 8)), ET, 0, c->stream>>>(m, k->n, A, k->n, x, beta, y, alpha);
  HB_LAUNCHED();
  return HB_OK;
}

// x = (B_k + D_x)^{-1} rhs
int hess_solve(hb_lowrank* k, const double* rhs, double* x)
{
  hb_ctx* c = k->ctx;
  if(k->n == 0) return HB_OK;
  if(k->l > 0) {
    HB_CHECK(multidot(k, k->DhInv, rhs, k->sigma)); // [sigma*S*(DhInv rhs); Y*(DhInv rhs)]
    HB_CHECK(hb_dense_sytrs(c, 2 * k->l, k->V, 2 * k->l, k->ipivV, k->p2l, 2 * k->l, 1));
  }
  k_lowrank_apply<<<stream_grid(c, k->n), ET, sizeof(double) * 2 * (k->l > 0 ? k->l : 1), c->stream>>>(
      k->n, k->l, k->sigma, k->St, k->Yt, k->n, k->p2l, k->DhInv, rhs, 0.0, nullptr, 0.0, 1.0, x);
  HB_LAUNCHED();
  return HB_OK;
}

int condense_enqueue(hb_lowrank* k, int mode, const double* fuse_rx = nullptr);
int condense_finish(hb_lowrank* k);

int resolve_global_n(hb_lowrank* k)
{
  hb_ctx* c = k->ctx;
  if(k->n_global >= 0) return HB_OK;
  if(c->nranks == 1) { k->n_global = k->n; return HB_OK; }
  const double nl = (double)k->n;
  HB_CUDA(cudaMemcpyAsync(k->stats + 2, &nl, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HB_CHECK(hb_allreduce_sum(c, k->stats + 2, 1));
  HB_CUDA(cudaMemcpyAsync(k->stats_host + 2, k->stats + 2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  k->n_global = (long long)(k->stats_host[2] + 0.5);
  return HB_OK;
}

// Enqueues the whole condensation (C_aug, all-reduce, V, N, equilibrated Cholesky) without touching the host. The info words are
// copied to pinned memory; condense_check() looks at them after a stream synchronisation.
int do_condense_async(hb_lowrank* k, const double* fuse_rx = nullptr)
{
  HB_REQUIRE(k->have_update, "hb_lowrank_condense: call hb_lowrank_update first");
  HB_REQUIRE(k->J || k->m == 0, "hb_lowrank_condense: Jacobian not set");
  const int Ma = k->m + 2 * k->l;
  HB_CHECK(refresh_rowptr(k));
  int mode = k->condense_mode;
  if(mode < 0) {
    // small systems: slicing + TMA setup do not pay off. The rule looks at the GLOBAL column count so that every rank (and every
    // world size) takes the same kernel for the same problem.
    HB_CHECK(resolve_global_n(k));
    mode = (k->n_global >= 32768 && Ma >= 64) ? 8 : 0;
  }
  HB_CHECK(condense_enqueue(k, mode, fuse_rx));
  k->check_pending = true;
  k->cond_valid = true; // optimistic: a failure is reported by the next synchronous call (hb_lowrank_check / hb_lowrank_condense)
  return HB_OK;
}

int condense_check(hb_lowrank* k)
{
  hb_ctx* c = k->ctx;
  if(!k->check_pending) return HB_OK;
  HB_CUDA(cudaStreamSynchronize(c->stream));
  k->check_pending = false;
  if(k->info_host[0] != 0) {
    k->cond_valid = false;
    return hb_fail(HB_ERR_NUMERIC, "hb_lowrank_condense: V is singular (BFGS inner matrix)%s", "");
  }
  if(k->info_host[1] != 0) {
    k->cond_valid = false;
    snprintf(g_hb_err, sizeof(g_hb_err), "hb_lowrank_condense: condensed matrix N is not SPD (leading minor %d)", k->info_host[1]);
    return HB_ERR_NUMERIC;
  }
  return HB_OK;
}

// synchronous condensation: reports breakdowns now; an int8-slice condensation chosen by the AUTO rule whose Cholesky breaks down is
// redone once with the exact FP64 kernel (the 5e-14 |N| perturbation of the emulation can cost positive definiteness of a nearly
// singular N late in the interior-point iteration; the reference's DPOSVX sees exact FP64 sums)
int do_condense(hb_lowrank* k)
{
  HB_CHECK(do_condense_async(k));
  int rc = condense_check(k);
  if(rc == HB_ERR_NUMERIC && k->condense_mode < 0 && k->condense_used != 0 && k->info_host[0] == 0) {
    HB_CHECK(condense_enqueue(k, 0));
    k->check_pending = true;
    k->cond_valid = true;
    k->fallbacks++;
    rc = condense_check(k);
  }
  return rc;
}

// fuse_rx (optional): the x-block of the right-hand side the caller is about to solve for. The int8-slice condensation has to sweep all
// rows for their maxima anyway; the same sweep then leaves tdot = [J; S; Y] (DhInv .* rx), from which step 2 of solveCompressed follows
// without reading J again (see solve_compressed).
int condense_enqueue(hb_lowrank* k, int mode, const double* fuse_rx)
{
  hb_ctx* c = k->ctx;
  const int m = k->m, l = k->l, Ma = m + 2 * l;
  k->tdot_valid = false;
  if(Ma > 0) {
    k->condense_used = mode;
    if(mode == 0) HB_CHECK(hb_syrk_rows(c, Ma, k->n, k->rowptr_dev, k->rows_aligned, k->DhInv, k->Caug, Ma));
    else {
      const bool fuse = fuse_rx && m > 0;
      if(fuse && k->n == 0) HB_CUDA(cudaMemsetAsync(k->tdot, 0, sizeof(double) * Ma, c->stream));
      HB_CHECK(hb_syrk_rows_ozaki(c, Ma, k->n, k->rowptr_dev, k->rows_aligned, k->DhInv, k->Caug, Ma, mode, fuse ? fuse_rx : nullptr, fuse ? k->tdot : nullptr));
      k->tdot_valid = fuse;
    }
  }
  return condense_finish(k);
}

// everything after C_aug = [J;S;Y] DhInv [J;S;Y]^T (local columns) is in k->Caug
int condense_finish(hb_lowrank* k)
{
  hb_ctx* c = k->ctx;
  const int m = k->m, l = k->l, Ma = m + 2 * l;
  HB_CUDA(cudaMemsetAsync(k->info, 0, sizeof(int) * 4, c->stream));
  hb_phase_mark(c, HB_PH_CAUG);
  if(Ma > 0 && c->nranks > 1) {
    // the symmetric C_aug travels as its packed upper triangle: Ma(Ma+1)/2 doubles instead of Ma^2
    const long long tot = (long long)Ma * (Ma + 1) / 2;
    if(!k->tri) HB_CHECK(dmalloc(&k->tri, (size_t)(k->m + 2 * k->lmax) * (k->m + 2 * k->lmax + 1) / 2 + (size_t)(k->m + 2 * k->lmax)));
    const int g = (int)((tot + 255) / 256 < (long long)c->num_sms * 8 ? (tot + 255) / 256 : (long long)c->num_sms * 8);
    k_pack_upper<<<g, 256, 0, c->stream>>>(Ma, k->Caug, Ma, k->tri);
    HB_LAUNCHED();
    // the fused row dots ride behind the triangle in the same reduction
    if(k->tdot_valid) HB_CUDA(cudaMemcpyAsync(k->tri + tot, k->tdot, sizeof(double) * Ma, cudaMemcpyDeviceToDevice, c->stream));
    HB_CHECK(hb_allreduce_sum(c, k->tri, tot + (k->tdot_valid ? Ma : 0)));
    k_unpack_upper<<<g, 256, 0, c->stream>>>(Ma, k->Caug, Ma, k->tri);
    HB_LAUNCHED();
    if(k->tdot_valid) HB_CUDA(cudaMemcpyAsync(k->tdot, k->tri + tot, sizeof(double) * Ma, cudaMemcpyDeviceToDevice, c->stream));
  }
  hb_phase_mark(c, HB_PH_ALLREDUCE);
  if(l > 0) {
    k_build_V<<<(4 * l * l + 127) / 128, 128, 0, c->stream>>>(m, l, k->sigma, k->Caug, Ma, k->SSt, k->Ld, k->Dd_sec, k->V);
    HB_LAUNCHED();
    HB_CHECK(hb_dense_sytf2(c, 2 * l, k->V, 2 * l, k->ipivV, k->info + 0));
    if(m > 0) {
      k_build_U<<<(m * 2 * l + 127) / 128, 128, 0, c->stream>>>(m, l, k->sigma, k->Caug, Ma, k->U, k->Z);
      HB_LAUNCHED();
      HB_CHECK(hb_dense_sytrs(c, 2 * l, k->V, 2 * l, k->ipivV, k->Z, 2 * l, m));
    }
  }
  if(m > 0) {
    const long long tot = (long long)m * m;
    k_form_N<<<(int)((tot + 255) / 256 < (long long)c->num_sms * 8 ? (tot + 255) / 256 : (long long)c->num_sms * 8), 256, 0, c->stream>>>(
        m, k->meq, l, k->Caug, Ma, k->U, k->Z, k->Dd_inv, k->Nmat);
    HB_LAUNCHED();
    HB_CHECK(hb_dense_equilibrate(c, m, k->Nmat, m, k->F, m, k->svec));
    hb_phase_mark(c, HB_PH_VN);
    HB_CHECK(hb_dense_chol_with_inverses(c, m, k->F, m, k->info + 1, k->Finv, &k->have_finv));
    hb_phase_mark(c, HB_PH_CHOL);
  }
  HB_CUDA(cudaMemcpyAsync(k->info_host, k->info, sizeof(int) * 4, cudaMemcpyDeviceToHost, c->stream));
  return HB_OK;
}

} // namespace

int hb_lr_gemv_rows(hb_lowrank* k, const double* A, int m, double beta, double* y, double alpha, const double* x)
{
  return gemv_rows(k, A, m, beta, y, alpha, x);
}
int hb_lr_gemv_cols(hb_lowrank* k, const double* A, int m, double beta, double* y, double alpha, const double* x)
{
  return gemv_cols(k, A, m, beta, y, alpha, x);
}
int hb_lr_multidot(hb_lowrank* k, const double* w, const double* x, double sigma_s) { return multidot(k, w, x, sigma_s); }
int hb_lr_refresh_rowptr(hb_lowrank* k) { return refresh_rowptr(k); }
int hb_lr_global_n(hb_lowrank* k, long long* n_global)
{
  HB_CHECK(resolve_global_n(k));
  *n_global = k->n_global;
  return HB_OK;
}

extern "C" int hb_lowrank_create(hb_ctx* c, long long n_local, int m_eq, int m_ineq, int l_max, hb_lowrank** out)
{
  HB_REQUIRE(c && out && n_local >= 0 && m_eq >= 0 && m_ineq >= 0 && l_max >= 0 && l_max <= 256, "hb_lowrank_create: bad arguments");
  HB_CUDA(cudaSetDevice(c->device));
  hb_lowrank* k = new hb_lowrank;
  k->ctx = c; k->n = n_local; k->meq = m_eq; k->mineq = m_ineq; k->m = m_eq + m_ineq; k->lmax = l_max;
  if(const char* e = getenv("HB_CONDENSE")) { // "oz6" | "oz7" | "oz8" | "dmma"
    if(e[0] == 'o' && e[1] == 'z' && e[2] >= '6' && e[2] <= '8') k->condense_mode = e[2] - '0';
    else if(e[0] == 'd') k->condense_mode = 0;
  }
  const int m = k->m, Mamax = m + 2 * l_max, l2 = 2 * l_max;
  HB_CHECK(dmalloc(&k->Dx, n_local)); HB_CHECK(dmalloc(&k->DhInv, n_local));
  HB_CHECK(dmalloc(&k->Dd, m_ineq)); HB_CHECK(dmalloc(&k->Dd_inv, m_ineq));
  HB_CHECK(dmalloc(&k->Caug, (size_t)Mamax * Mamax));
  HB_CHECK(dmalloc(&k->SSt, (size_t)l_max * l_max)); HB_CHECK(dmalloc(&k->Ld, (size_t)l_max * l_max)); HB_CHECK(dmalloc(&k->Dd_sec, l_max));
  HB_CHECK(dmalloc(&k->V, (size_t)l2 * l2)); HB_CHECK(dmalloc(&k->Mdir, (size_t)l2 * l2));
  HB_CHECK(dmalloc(&k->U, (size_t)m * l2)); HB_CHECK(dmalloc(&k->Z, (size_t)m * l2));
  HB_CHECK(dmalloc(&k->Nmat, (size_t)m * m)); HB_CHECK(dmalloc(&k->F, (size_t)m * m));
  HB_CHECK(dmalloc(&k->svec, m)); HB_CHECK(dmalloc(&k->rhs, m)); HB_CHECK(dmalloc(&k->dy, m)); HB_CHECK(dmalloc(&k->work, 2 * (size_t)m + 2));
  HB_CHECK(dmalloc(&k->Finv, HB_CHOL_INV_DOUBLES(m > 0 ? m : 1)));
  HB_CHECK(dmalloc(&k->stats, 4));
  HB_CHECK(dmalloc(&k->nv1, n_local)); HB_CHECK(dmalloc(&k->nv2, n_local));
  HB_CHECK(dmalloc(&k->tdot, (size_t)Mamax));
  HB_CHECK(dmalloc(&k->p2l, l2));
  k->md_grid = stream_grid(c, n_local);
  HB_CHECK(dmalloc(&k->md_partial, (size_t)k->md_grid * (l2 > 0 ? l2 : 1)));
  HB_CHECK(dmalloc(&k->mi1, m_ineq)); HB_CHECK(dmalloc(&k->mi2, m_ineq)); HB_CHECK(dmalloc(&k->mi3, m_ineq));
  HB_CUDA(cudaMalloc(&k->ipivV, sizeof(int) * (l2 + 1))); HB_CUDA(cudaMalloc(&k->ipivM, sizeof(int) * (l2 + 1)));
  HB_CUDA(cudaMalloc(&k->info, sizeof(int) * 4));
  HB_CUDA(cudaMalloc(&k->rowptr_dev, sizeof(double*) * (Mamax + 2)));
  HB_CUDA(cudaMallocHost(&k->rowptr_host, sizeof(double*) * (Mamax + 2)));
  HB_CUDA(cudaMallocHost(&k->info_host, sizeof(int) * 4));
  HB_CUDA(cudaMallocHost(&k->stats_host, sizeof(double) * 4));
  *out = k;
  return HB_OK;
}

extern "C" int hb_lowrank_destroy(hb_lowrank* k)
{
  if(!k) return HB_OK;
  cudaSetDevice(k->ctx->device);
  cudaStreamSynchronize(k->ctx->stream);
  double* bufs[] = {k->Dx, k->DhInv, k->Dd, k->Dd_inv, k->Jpack, k->Caug, k->SSt, k->Ld, k->Dd_sec, k->V, k->Mdir, k->U, k->Z, k->Nmat, k->F,
                    k->svec, k->rhs, k->dy, k->work, k->stats, k->nv1, k->nv2, k->p2l, k->md_partial, k->mi1, k->mi2, k->mi3, k->hJ, k->kry, k->kry_m, k->sec_S, k->sec_Y, k->sec_xprev, k->sec_gprev, k->sec_Jprev, k->lsq_M, k->Finv, k->Ctmp, k->tri, k->tdot};
  for(double* b : bufs) if(b) cudaFree(b);
  for(double* b : k->hbuf) if(b) cudaFree(b);
  cudaFree(k->ipivV); cudaFree(k->ipivM); cudaFree(k->info); cudaFree(k->rowptr_dev);
  cudaFreeHost(k->rowptr_host); cudaFreeHost(k->info_host); cudaFreeHost(k->stats_host);
  if(k->copy_stream) {
    cudaStreamDestroy(k->copy_stream);
    for(cudaEvent_t e : k->chunk_ev) if(e) cudaEventDestroy(e);
    cudaFree(k->chunk_rowptr_dev); cudaFreeHost(k->chunk_rowptr_host);
  }
  delete k;
  return HB_OK;
}

extern "C" int hb_lowrank_set_patterns(hb_lowrank* k, const double* ixl, const double* ixu, const double* idl, const double* idu)
{
  HB_REQUIRE(k, "null handle");
  HB_REQUIRE((ixl && ixu) || k->n == 0, "hb_lowrank_set_patterns: null x pattern");
  HB_REQUIRE((idl && idu) || k->mineq == 0, "hb_lowrank_set_patterns: null d pattern");
  k->ixl = ixl; k->ixu = ixu; k->idl = idl; k->idu = idu;
  k->cond_valid = false;
  return HB_OK;
}

extern "C" int hb_lowrank_set_jacobian(hb_lowrank* k, const double* Jc, const double* Jd)
{
  HB_REQUIRE(k, "null handle");
  HB_REQUIRE((Jc || k->meq == 0) && (Jd || k->mineq == 0), "hb_lowrank_set_jacobian: null Jacobian");
  hb_ctx* c = k->ctx;
  const double* J;
  if(k->meq == 0) J = Jd;
  else if(k->mineq == 0 || Jd == Jc + (size_t)k->meq * k->n) J = Jc;
  else {
    if(!k->Jpack) HB_CHECK(dmalloc(&k->Jpack, (size_t)k->m * k->n));
    HB_CUDA(cudaMemcpyAsync(k->Jpack, Jc, sizeof(double) * (size_t)k->meq * k->n, cudaMemcpyDeviceToDevice, c->stream));
    HB_CUDA(cudaMemcpyAsync(k->Jpack + (size_t)k->meq * k->n, Jd, sizeof(double) * (size_t)k->mineq * k->n, cudaMemcpyDeviceToDevice, c->stream));
    J = k->Jpack;
  }
  if(J != k->J) k->rowptr_dirty = true;
  k->J = J;
  k->cond_valid = false;
  return HB_OK;
}

extern "C" int hb_lowrank_set_secant(hb_lowrank* k, int l, double sigma, const double* St, const double* Yt, const double* L_host,
                                     const double* D_host)
{
  HB_REQUIRE(k && l >= 0 && l <= k->lmax, "hb_lowrank_set_secant: bad memory length");
  HB_REQUIRE(l == 0 || (St && Yt && L_host && D_host), "hb_lowrank_set_secant: null argument");
  hb_ctx* c = k->ctx;
  if(l != k->l || St != k->St || Yt != k->Yt) k->rowptr_dirty = true;
  k->l = l; k->sigma = sigma; k->St = St; k->Yt = Yt;
  k->cond_valid = false;
  k->mdir_valid = false;
  k->have_update = false; // DhInv depends on sigma
  if(l > 0) {
    HB_CUDA(cudaMemcpyAsync(k->Ld, L_host, sizeof(double) * l * l, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaMemcpyAsync(k->Dd_sec, D_host, sizeof(double) * l, cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream)); // L_host / D_host are caller-owned pageable memory
    // S S^T (l x l) -- depends only on the secant memory, not on the barrier diagonal
    for(int q = 0; q < l; q++) k->rowptr_host[q] = St + (size_t)q * k->n;
    bool al = true;
    for(int q = 0; q < l; q++) al = al && ((reinterpret_cast<uintptr_t>(k->rowptr_host[q]) & 15u) == 0);
    HB_CUDA(cudaMemcpyAsync(k->rowptr_dev, k->rowptr_host, sizeof(double*) * l, cudaMemcpyHostToDevice, c->stream));
    HB_CHECK(hb_syrk_rows(c, l, k->n, k->rowptr_dev, al, nullptr, k->SSt, l));
    HB_CHECK(hb_allreduce_sum(c, k->SSt, (long long)l * l));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    k->rowptr_dirty = true;
  }
  return HB_OK;
}

extern "C" int hb_lowrank_update(hb_lowrank* k, const double* zl, const double* sxl, const double* zu, const double* sxu, const double* vl,
                                 const double* sdl, const double* vu, const double* sdu)
{
  HB_REQUIRE(k, "null handle");
  HB_REQUIRE(k->n == 0 || (zl && sxl && zu && sxu), "hb_lowrank_update: null x-side iterate block");
  HB_REQUIRE(k->mineq == 0 || (vl && sdl && vu && sdu), "hb_lowrank_update: null d-side iterate block");
  HB_REQUIRE(k->n == 0 || k->ixl, "hb_lowrank_update: patterns not set");
  hb_ctx* c = k->ctx;
  k->zl = zl; k->sxl = sxl; k->zu = zu; k->sxu = sxu; k->vl = vl; k->sdl = sdl; k->vu = vu; k->sdu = sdu;
  hb_phase_mark(c, HB_PH_START);
  if(k->n > 0) {
    k_update_x<<<stream_grid(c, k->n), ET, 0, c->stream>>>(k->n, zl, sxl, zu, sxu, k->ixl, k->ixu, k->sigma, k->Dx, k->DhInv);
    HB_LAUNCHED();
  }
  if(k->mineq > 0) {
    k_update_d<<<(k->mineq + 127) / 128, 128, 0, c->stream>>>(k->mineq, vl, sdl, vu, sdu, k->idl, k->idu, k->Dd, k->Dd_inv);
    HB_LAUNCHED();
  }
  hb_phase_mark(c, HB_PH_UPDATE);
  k->have_update = true;
  k->cond_valid = false;
  return HB_OK;
}

extern "C" int hb_lowrank_set_condense_mode(hb_lowrank* k, int mode)
{
  HB_REQUIRE(k && (mode == -1 || mode == 0 || mode == 6 || mode == 7 || mode == 8), "hb_lowrank_set_condense_mode: mode must be -1, 0, 6, 7 or 8");
  k->condense_mode = mode;
  k->cond_valid = false;
  return HB_OK;
}

extern "C" int hb_lowrank_get_condense_mode(hb_lowrank* k) { return k ? k->condense_used : 0; }

extern "C" int hb_lowrank_condense(hb_lowrank* k)
{
  HB_REQUIRE(k, "null handle");
  return do_condense(k);
}

extern "C" int hb_lowrank_condense_async(hb_lowrank* k)
{
  HB_REQUIRE(k, "null handle");
  return do_condense_async(k);
}
extern "C" int hb_lowrank_check(hb_lowrank* k)
{
  HB_REQUIRE(k, "null handle");
  return condense_check(k);
}
extern "C" int hb_lowrank_fallback_count(hb_lowrank* k) { return k ? k->fallbacks : 0; }

extern "C" int hb_lowrank_hess_solve(hb_lowrank* k, const double* rhs, double* x)
{
  HB_REQUIRE(k && (k->n == 0 || (rhs && x)), "hb_lowrank_hess_solve: null argument");
  if(!k->cond_valid) HB_CHECK(do_condense_async(k));
  return hess_solve(k, rhs, x);
}

extern "C" int hb_lowrank_solve_compressed(hb_lowrank* k, double* rx, const double* ryc, const double* ryd, double* dx, double* dyc, double* dyd)
{
  HB_REQUIRE(k, "null handle");
  HB_REQUIRE(k->n == 0 || (rx && dx), "hb_lowrank_solve_compressed: null x block");
  HB_REQUIRE((k->meq == 0 || (ryc && dyc)) && (k->mineq == 0 || (ryd && dyd)), "hb_lowrank_solve_compressed: null dual block");
  hb_ctx* c = k->ctx;
  const int m = k->m;
  // A pending condensation is enqueued here (breakdowns surface at the next synchronous call, hb_lowrank_check). With the int8-slice
  // kernel its row-maximum sweep over [J; S; Y] also produces tdot = [J; S; Y] (DhInv .* rx), and steps 1-2 collapse to
  //   J (H+Dx)^{-1} rx = tdot_J - Z [sigma*tdot_S; tdot_Y],   Z = U V^{-1}  (U = [sigma J DhInv S^T, J DhInv Y^T], kept from the condensation)
  // which is the same product with the low-rank correction applied on the m side: J is not read a second time.
  bool fused = false;
  if(!k->cond_valid) {
    HB_CHECK(do_condense_async(k, rx));
    fused = k->tdot_valid;
    k->tdot_valid = false; // tied to this rx
  }
  if(fused) {
    hb_phase_mark(c, HB_PH_HSOLVE1);
    k_fused_rhs<<<(m + 127) / 128, 128, 0, c->stream>>>(m, k->meq, k->l, k->sigma, k->tdot, k->Z, ryc, ryd, k->rhs);
    HB_LAUNCHED();
    hb_phase_mark(c, HB_PH_JX);
  } else {
    // 1. dx_tmp = (H+Dx)^{-1} rx                                  hiopKKTLinSys.cpp:1146
    HB_CHECK(hess_solve(k, rx, dx));
    hb_phase_mark(c, HB_PH_HSOLVE1);
    if(m > 0) {
      // 2. rhs = J*dx_tmp - [ryc; ryd]                              :1154-1157
      HB_CHECK(gemv_rows(k, k->J, m, 0.0, k->rhs, 1.0, dx));
      hb_phase_mark(c, HB_PH_JX);
      k_sub_stacked<<<(m + 127) / 128, 128, 0, c->stream>>>(k->meq, k->mineq, k->rhs, ryc, ryd);
      HB_LAUNCHED();
    }
  }
  if(m > 0) {
    // 3. N dy = rhs with residual-driven refinement               :1169, 1192-1350
    HB_CHECK(hb_dense_spd_solve_refine2(c, m, k->F, m, k->have_finv ? k->Finv : nullptr, k->svec, k->Nmat, m, k->rhs, k->dy, k->work, 1e-8, 3,
                                        k->stats));
    hb_phase_mark(c, HB_PH_SPDSOLVE);
    if(k->meq) HB_CUDA(cudaMemcpyAsync(dyc, k->dy, sizeof(double) * k->meq, cudaMemcpyDeviceToDevice, c->stream));
    if(k->mineq) HB_CUDA(cudaMemcpyAsync(dyd, k->dy + k->meq, sizeof(double) * k->mineq, cudaMemcpyDeviceToDevice, c->stream));
    // 4. rx = rx - J^T dy                                          :1178
    HB_CHECK(gemv_cols(k, k->J, m, 1.0, rx, -1.0, k->dy));
    hb_phase_mark(c, HB_PH_JTY);
    HB_CUDA(cudaMemcpyAsync(k->stats_host, k->stats, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
  }
  // 5. dx = (H+Dx)^{-1} rx                                        :1180
  HB_CHECK(hess_solve(k, rx, dx));
  hb_phase_mark(c, HB_PH_HSOLVE2);
  return HB_OK;
}

extern "C" int hb_lowrank_last_solve_stats(hb_lowrank* k, int* n_refine, double* resid_inf)
{
  HB_REQUIRE(k, "null handle");
  HB_CHECK(condense_check(k));
  HB_CUDA(cudaStreamSynchronize(k->ctx->stream));
  if(n_refine) *n_refine = k->m > 0 ? (int)k->stats_host[0] : 0;
  if(resid_inf) *resid_inf = k->m > 0 ? k->stats_host[1] : 0.0;
  return HB_OK;
}

extern "C" int hb_lowrank_compute_directions(hb_lowrank* k, const double* const* res, double* const* dir)
{
  HB_REQUIRE(k && res && dir, "hb_lowrank_compute_directions: null argument");
  HB_REQUIRE(k->have_update, "hb_lowrank_compute_directions: call hb_lowrank_update first");
  hb_ctx* c = k->ctx;
  enum { RX, RD, RYC, RYD, RXL, RXU, RDL, RDU, RSZL, RSZU, RSVL, RSVU };
  enum { DX, DD, DYC, DYD, DSXL, DSXU, DSDL, DSDU, DZL, DZU, DVL, DVU };
  const long long n = k->n;
  const int mi = k->mineq;
  double* rx_tilde = k->nv1;
  double* ryd2 = k->mi1;
  double* ryd_tilde = k->mi2;
  if(n > 0) {
    k_reduce_rhs<<<stream_grid(c, n), ET, 0, c->stream>>>(n, res[RX], res[RSZL], k->zl, res[RXL], k->sxl, k->ixl, res[RSZU], k->zu, res[RXU], k->sxu,
                                                          k->ixu, rx_tilde);
    HB_LAUNCHED();
  }
  if(mi > 0) {
    k_reduce_rhs<<<stream_grid(c, mi), ET, 0, c->stream>>>(mi, res[RD], res[RSVL], k->vl, res[RDL], k->sdl, k->idl, res[RSVU], k->vu, res[RDU],
                                                           k->sdu, k->idu, ryd2);
    HB_LAUNCHED();
    k_axzpy_small<<<(mi + 127) / 128, 128, 0, c->stream>>>(mi, ryd_tilde, res[RYD], ryd2, k->Dd_inv);
    HB_LAUNCHED();
  }
  HB_CHECK(hb_lowrank_solve_compressed(k, rx_tilde, res[RYC], ryd_tilde, dir[DX], dir[DYC], dir[DYD]));
  if(mi > 0) {
    k_recover_dd<<<(mi + 127) / 128, 128, 0, c->stream>>>(mi, dir[DD], ryd2, dir[DYD], k->Dd_inv);
    HB_LAUNCHED();
    k_recover_slack_duals<<<stream_grid(c, mi), ET, 0, c->stream>>>(mi, dir[DD], res[RDL], res[RSVL], k->vl, k->sdl, k->idl, res[RDU], res[RSVU],
                                                                    k->vu, k->sdu, k->idu, dir[DSDL], dir[DVL], dir[DSDU], dir[DVU]);
    HB_LAUNCHED();
  }
  if(n > 0) {
    k_recover_slack_duals<<<stream_grid(c, n), ET, 0, c->stream>>>(n, dir[DX], res[RXL], res[RSZL], k->zl, k->sxl, k->ixl, res[RXU], res[RSZU], k->zu,
                                                                   k->sxu, k->ixu, dir[DSXL], dir[DZL], dir[DSXU], dir[DZU]);
    HB_LAUNCHED();
  }
  return HB_OK;
}

extern "C" int hb_lowrank_hess_times_vec(hb_lowrank* k, double beta, double* y, double alpha, const double* x, int add_log_term)
{
  HB_REQUIRE(k && (k->n == 0 || (x && y)), "hb_lowrank_hess_times_vec: null argument");
  HB_REQUIRE(!add_log_term || k->have_update, "hb_lowrank_hess_times_vec: Dx not available (call update)");
  hb_ctx* c = k->ctx;
  const int l = k->l;
  if(k->n == 0) return HB_OK;
  if(l > 0) {
    if(!k->mdir_valid) {
      HB_CUDA(cudaMemsetAsync(k->info + 2, 0, sizeof(int), c->stream));
      k_build_Mdirect<<<(4 * l * l + 127) / 128, 128, 0, c->stream>>>(l, k->sigma, k->SSt, k->Ld, k->Dd_sec, k->Mdir);
      HB_LAUNCHED();
      HB_CHECK(hb_dense_sytf2(c, 2 * l, k->Mdir, 2 * l, k->ipivM, k->info + 2));
      k->mdir_valid = true;
    }
    HB_CHECK(multidot(k, nullptr, x, k->sigma)); // [sigma S x; Y x]
    HB_CHECK(hb_dense_sytrs(c, 2 * l, k->Mdir, 2 * l, k->ipivM, k->p2l, 2 * l, 1));
  }
  k_lowrank_apply<<<stream_grid(c, k->n), ET, sizeof(double) * 2 * (l > 0 ? l : 1), c->stream>>>(
      k->n, l, k->sigma, k->St, k->Yt, k->n, k->p2l, nullptr, x, k->sigma, add_log_term ? k->Dx : nullptr, beta, alpha, y);
  HB_LAUNCHED();
  return HB_OK;
}

extern "C" const double* hb_lowrank_Dx(hb_lowrank* k) { return k ? k->Dx : nullptr; }
extern "C" const double* hb_lowrank_DhInv(hb_lowrank* k) { return k ? k->DhInv : nullptr; }
extern "C" const double* hb_lowrank_Dd_inv(hb_lowrank* k) { return k ? k->Dd_inv : nullptr; }
extern "C" const double* hb_lowrank_N(hb_lowrank* k) { return k ? k->Nmat : nullptr; }

// ---- one whole KKT system from host buffers ----------------------------------------------------------------------------
extern "C" int hb_lowrank_kkt_system_host(hb_lowrank* k, const double* Jc_host, const double* Jd_host, const double* zl, const double* sxl,
                                          const double* zu, const double* sxu, const double* vl, const double* sdl, const double* vu,
                                          const double* sdu, const double* rx, const double* ryc, const double* ryd, double* dx, double* dyc,
                                          double* dyd)
{
  HB_REQUIRE(k, "null handle");
  hb_ctx* c = k->ctx;
  const long long n = k->n;
  const int meq = k->meq, mi = k->mineq;
  // device staging: 0..3 x-side iterate, 4..7 d-side iterate, 8 rx, 9 ryc, 10 ryd, 11 dx, 12 dyc, 13 dyd
  const size_t sz[14] = {(size_t)n, (size_t)n, (size_t)n, (size_t)n, (size_t)mi, (size_t)mi, (size_t)mi, (size_t)mi, (size_t)n, (size_t)meq, (size_t)mi,
                         (size_t)n, (size_t)meq, (size_t)mi};
  for(int i = 0; i < 14; i++)
    if(!k->hbuf[i]) HB_CHECK(dmalloc(&k->hbuf[i], sz[i]));
  const double* src[11] = {zl, sxl, zu, sxu, vl, sdl, vu, sdu, rx, ryc, ryd};
  for(int i = 0; i < 11; i++)
    if(sz[i]) {
      HB_REQUIRE(src[i], "hb_lowrank_kkt_system_host: null host input");
      HB_CUDA(cudaMemcpyAsync(k->hbuf[i], src[i], sizeof(double) * sz[i], cudaMemcpyHostToDevice, c->stream));
    }
  const int m = k->m, Ma = m + 2 * k->l;
  const bool have_J = Jc_host || Jd_host;
  // The 8 m n bytes of J dominate this call (PCIe). When they are large, J is uploaded in column chunks on a second stream and
  // each chunk is condensed (exact FP64 DMMA kernel, which needs no global row scaling) while the next one is in flight; the
  // partial C_aug are added in chunk order, so the result does not depend on timing.
  static const size_t chunk_min_bytes = getenv("HB_HOST_CHUNK_MIN_BYTES") ? (size_t)atoll(getenv("HB_HOST_CHUNK_MIN_BYTES")) : ((size_t)256 << 20);
  const bool chunked = have_J && Ma > 0 && (k->condense_mode <= 0) && (size_t)m * n * sizeof(double) >= chunk_min_bytes && n >= 2048;
  if(have_J && !k->hJ) HB_CHECK(dmalloc(&k->hJ, (size_t)k->m * n));
  if(have_J && !chunked) {
    if(meq) HB_CUDA(cudaMemcpyAsync(k->hJ, Jc_host, sizeof(double) * (size_t)meq * n, cudaMemcpyHostToDevice, c->stream));
    if(mi) HB_CUDA(cudaMemcpyAsync(k->hJ + (size_t)meq * n, Jd_host, sizeof(double) * (size_t)mi * n, cudaMemcpyHostToDevice, c->stream));
  }
  if(have_J) HB_CHECK(hb_lowrank_set_jacobian(k, k->hJ, k->hJ + (size_t)meq * n));
  HB_CHECK(hb_lowrank_update(k, k->hbuf[0], k->hbuf[1], k->hbuf[2], k->hbuf[3], k->hbuf[4], k->hbuf[5], k->hbuf[6], k->hbuf[7]));
  if(!chunked) {
    HB_CHECK(do_condense(k));
  } else {
    constexpr int NCH = 16;
    if(!k->copy_stream) {
      HB_CUDA(cudaStreamCreateWithFlags(&k->copy_stream, cudaStreamNonBlocking));
      for(int q = 0; q < 32; q++) HB_CUDA(cudaEventCreateWithFlags(&k->chunk_ev[q], cudaEventDisableTiming));
      HB_CUDA(cudaMalloc(&k->chunk_rowptr_dev, sizeof(double*) * 32 * (size_t)(k->m + 2 * k->lmax)));
      HB_CUDA(cudaMallocHost(&k->chunk_rowptr_host, sizeof(double*) * 32 * (size_t)(k->m + 2 * k->lmax)));
    }
    if(!k->Ctmp) HB_CHECK(dmalloc(&k->Ctmp, (size_t)(k->m + 2 * k->lmax) * (k->m + 2 * k->lmax)));
    HB_CHECK(refresh_rowptr(k)); // k->rowptr_host: full-length rows of [J; S; Y]
    long long csz = ((n + NCH - 1) / NCH + 63) & ~63LL;
    int nch = (int)((n + csz - 1) / csz);
    for(int q = 0; q < nch; q++)
      for(int i = 0; i < Ma; i++) k->chunk_rowptr_host[(size_t)q * Ma + i] = k->rowptr_host[i] + q * csz;
    HB_CUDA(cudaMemcpyAsync(k->chunk_rowptr_dev, k->chunk_rowptr_host, sizeof(double*) * (size_t)nch * Ma, cudaMemcpyHostToDevice, c->stream));
    // the copy stream starts after the (small) uploads above were enqueued; it only ever writes k->hJ
    HB_CUDA(cudaEventRecord(k->chunk_ev[31], c->stream));
    HB_CUDA(cudaStreamWaitEvent(k->copy_stream, k->chunk_ev[31], 0));
    // Copies are submitted only LOOKAHEAD chunks ahead of the kernels that consume them: if the two streams ever share a hardware
    // work queue (CUDA_DEVICE_MAX_CONNECTIONS) commands run in submission order, and "all copies, then all kernels" would serialise.
    constexpr int LOOKAHEAD = 3;
    auto submit_copy = [&](int q) -> int {
      const long long c0 = q * csz, w = (c0 + csz <= n ? csz : n - c0);
      if(meq)
        HB_CUDA(cudaMemcpy2DAsync(k->hJ + c0, sizeof(double) * n, Jc_host + c0, sizeof(double) * n, sizeof(double) * w, meq, cudaMemcpyHostToDevice,
                                  k->copy_stream));
      if(mi)
        HB_CUDA(cudaMemcpy2DAsync(k->hJ + (size_t)meq * n + c0, sizeof(double) * n, Jd_host + c0, sizeof(double) * n, sizeof(double) * w, mi,
                                  cudaMemcpyHostToDevice, k->copy_stream));
      HB_CUDA(cudaEventRecord(k->chunk_ev[q], k->copy_stream));
      return HB_OK;
    };
    for(int q = 0; q < nch && q < LOOKAHEAD; q++) HB_CHECK(submit_copy(q));
    k->condense_used = 0;
    // timing-enabled cudaEventRecord between the kernels serialises the compute stream against the copy engine (measured: with
    // hb_ctx_enable_timing the first kernel started when the last copy ended, 183 ms instead of 151 ms per call) -> off in here
    const bool timing_saved = c->timing;
    c->timing = false;
    for(int q = 0; q < nch; q++) {
      const long long c0 = q * csz, w = (c0 + csz <= n ? csz : n - c0);
      HB_CUDA(cudaStreamWaitEvent(c->stream, k->chunk_ev[q], 0));
      // csz is a multiple of 64 columns: every chunk keeps the 16-byte alignment of its rows
      HB_CHECK(hb_syrk_rows(c, Ma, w, k->chunk_rowptr_dev + (size_t)q * Ma, k->rows_aligned, k->DhInv + c0, q == 0 ? k->Caug : k->Ctmp, Ma));
      if(q > 0) HB_CHECK(hb_vec_axpy(c, (long long)Ma * Ma, k->Caug, 1.0, k->Ctmp));
      if(q + LOOKAHEAD < nch) HB_CHECK(submit_copy(q + LOOKAHEAD));
    }
    c->timing = timing_saved;
    HB_CHECK(condense_finish(k));
    k->check_pending = true;
    k->cond_valid = true;
    HB_CHECK(condense_check(k));
  }
  HB_CHECK(hb_lowrank_solve_compressed(k, k->hbuf[8], k->hbuf[9], k->hbuf[10], k->hbuf[11], k->hbuf[12], k->hbuf[13]));
  if(n) HB_CUDA(cudaMemcpyAsync(dx, k->hbuf[11], sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
  if(meq) HB_CUDA(cudaMemcpyAsync(dyc, k->hbuf[12], sizeof(double) * meq, cudaMemcpyDeviceToHost, c->stream));
  if(mi) HB_CUDA(cudaMemcpyAsync(dyd, k->hbuf[13], sizeof(double) * mi, cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  return HB_OK;
}

// ---- public gemv (hiopMatrixDenseRowMajor::timesVec / transTimesVec) ----------------------------------------------------
extern "C" int hb_mat_times_vec(hb_ctx* c, int m, long long n, const double* A, long long lda, double beta, double* y, double alpha, const double* x)
{
  HB_REQUIRE(c && m >= 0 && n >= 0 && lda >= n, "hb_mat_times_vec: bad arguments");
  if(m == 0) return HB_OK;
  const int nchunks = (int)((n + GR_CHUNK - 1) / GR_CHUNK);
  HB_CHECK(hb_ws_reserve(c, sizeof(double) * (size_t)(nchunks > 0 ? nchunks : 1) * m));
  if(nchunks > 0) {
    k_gemv_rows_partial<<<nchunks, GR_THREADS, 0, c->stream>>>(m, n, A, lda, x, (double*)c->ws);
    HB_LAUNCHED();
  }
  const double b = (c->nranks > 1 && c->rank != 0) ? 0.0 : beta;
  k_gemv_rows_final<<<(m + 31) / 32, 256, 0, c->stream>>>(m, nchunks, (const double*)c->ws, b, y, alpha);
  HB_LAUNCHED();
  if(c->nranks > 1) HB_CHECK(hb_allreduce_sum(c, y, m));
  return HB_OK;
}
extern "C" int hb_mat_trans_times_vec(hb_ctx* c, int m, long long n, const double* A, long long lda, double beta, double* y, double alpha,
                                      const double* x)
{
  HB_REQUIRE(c && m >= 0 && n >= 0 && lda >= n, "hb_mat_trans_times_vec: bad arguments");
  if(n == 0) return HB_OK;
  const long long pairs = (n + 1) / 2;
  k_gemv_cols<1><<<(unsigned)((pairs + ET - 1) / ET), ET, 0, c->stream>>>(m, n, A, lda, x, beta, y, alpha);
  HB_LAUNCHED();
  return HB_OK;
}
