// hb_lowrank_kkt_system_host from a plain C++ process (no PyTorch): wall-clock per call.
// g++ -O2 -I../include e2e_probe.cpp -L../hiop_b200 -lhiopb200 -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/../hiop_b200 -o e2e_probe
#include "hiopb200.h"
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#define CK(x) do { int rc_ = (x); if(rc_ != 0) { fprintf(stderr, "%s failed: %s\n", #x, hb_last_error()); return 1; } } while(0)
int main(int argc, char** argv)
{
  const long long n = argc > 1 ? atoll(argv[1]) : 1000000;
  const int me = 500, mi = 500, l = 6, m = me + mi;
  hb_ctx* c; CK(hb_ctx_create(0, &c));
  hb_lowrank* k; CK(hb_lowrank_create(c, n, me, mi, l, &k));
  auto pinned = [](size_t cnt) { double* p; cudaHostAlloc(&p, sizeof(double) * cnt, cudaHostAllocDefault); return p; };
  double* J = pinned((size_t)m * n);
  unsigned long long s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; };
  const double sc = 1.0 / std::sqrt((double)n);
  for(size_t i = 0; i < (size_t)m * n; i++) J[i] = (rnd() - 0.5) * 3.4641 * sc;
  double *v[8], *rx = pinned(n), *ryc = pinned(me), *ryd = pinned(mi), *dx = pinned(n), *dyc = pinned(me), *dyd = pinned(mi);
  const size_t vs[8] = {(size_t)n, (size_t)n, (size_t)n, (size_t)n, (size_t)mi, (size_t)mi, (size_t)mi, (size_t)mi};
  for(int q = 0; q < 8; q++) { v[q] = pinned(vs[q]); for(size_t i = 0; i < vs[q]; i++) v[q][i] = 1e-3 + rnd(); }
  for(long long i = 0; i < n; i++) rx[i] = rnd() - 0.5;
  for(int i = 0; i < me; i++) ryc[i] = rnd() - 0.5;
  for(int i = 0; i < mi; i++) ryd[i] = rnd() - 0.5;
  // patterns (all lower-bounded), secant memory S, Y = S * U(0.5, 2)
  std::vector<double> ones(n, 1.0), St((size_t)l * n), Yt((size_t)l * n), L(l * l, 0.0), D(l, 0.0);
  for(size_t i = 0; i < St.size(); i++) { St[i] = rnd() - 0.5; Yt[i] = St[i] * (0.5 + 1.5 * rnd()); }
  for(int a = 0; a < l; a++) for(int b = 0; b <= a; b++) { double t = 0; for(long long i = 0; i < n; i++) t += St[a * n + i] * Yt[b * n + i]; if(a == b) D[a] = t; else L[a * l + b] = t; }
  double *dixl, *dixu, *didl, *didu, *dS, *dY;
  CK(hb_malloc(c, 8 * n, (void**)&dixl)); CK(hb_malloc(c, 8 * n, (void**)&dixu)); CK(hb_malloc(c, 8 * mi, (void**)&didl)); CK(hb_malloc(c, 8 * mi, (void**)&didu));
  CK(hb_malloc(c, 8 * (size_t)l * n, (void**)&dS)); CK(hb_malloc(c, 8 * (size_t)l * n, (void**)&dY));
  CK(hb_memcpy_h2d(c, dixl, ones.data(), 8 * n)); CK(hb_memcpy_h2d(c, dixu, ones.data(), 8 * n));
  CK(hb_memcpy_h2d(c, didl, ones.data(), 8 * mi)); CK(hb_memcpy_h2d(c, didu, ones.data(), 8 * mi));
  CK(hb_memcpy_h2d(c, dS, St.data(), 8 * (size_t)l * n)); CK(hb_memcpy_h2d(c, dY, Yt.data(), 8 * (size_t)l * n));
  CK(hb_ctx_sync(c));
  CK(hb_lowrank_set_patterns(k, dixl, dixu, didl, didu));
  CK(hb_lowrank_set_secant(k, l, 1.0, dS, dY, L.data(), D.data()));
  if(getenv("PROBE_TIMING")) hb_ctx_enable_timing(c, 1);
  for(int rep = 0; rep < 4; rep++) {
    auto t0 = std::chrono::steady_clock::now();
    CK(hb_lowrank_kkt_system_host(k, J, J + (size_t)me * n, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], rx, ryc, ryd, dx, dyc, dyd));
    auto t1 = std::chrono::steady_clock::now();
    int nref; double resid; hb_lowrank_last_solve_stats(k, &nref, &resid);
    printf("call %d: %.1f ms  (mode %d, refinements %d, residual %.2e, dx[0] %.6e)\n", rep, std::chrono::duration<double, std::milli>(t1 - t0).count(),
           hb_lowrank_get_condense_mode(k), nref, resid, dx[0]);
  }
  return 0;
}
