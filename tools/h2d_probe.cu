// H2D bandwidth: one contiguous copy vs column-chunked 2-D copies (m rows x n doubles, 16 chunks), with and without a concurrent kernel
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin_big(double* p, long long n, int iters) // one CTA per SM: 200 KB of dynamic shared memory
{
  extern __shared__ double sm[];
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0;
  for(int it = 0; it < iters; it++)
    for(long long k = i; k < n; k += (long long)gridDim.x * blockDim.x) v += p[k];
  sm[threadIdx.x] = v;
  if(v == 123.456) p[0] = sm[0];
}
__global__ void spin(double* p, long long n, int iters)
{
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0;
  for(int it = 0; it < iters; it++)
    for(long long k = i; k < n; k += (long long)gridDim.x * blockDim.x) v += p[k];
  if(v == 123.456) p[0] = v;
}
int main()
{
  const long long m = 1000, n = 1000000;
  double *h, *d, *scratch;
  cudaHostAlloc(&h, sizeof(double) * m * n, cudaHostAllocDefault);
  cudaMalloc(&d, sizeof(double) * m * n);
  cudaMalloc(&scratch, sizeof(double) * (1 << 28));
  for(long long i = 0; i < m * n; i += 4096) h[i] = 1.0;
  cudaStream_t cs, ks; cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&ks, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for(int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0, cs);
    cudaMemcpyAsync(d, h, sizeof(double) * m * n, cudaMemcpyHostToDevice, cs);
    cudaEventRecord(e1, cs); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("contiguous: %.1f ms = %.1f GB/s\n", ms, 8e-6 * m * n / ms);
    for(int nch : {16, 64}) {
      const long long csz = n / nch;
      cudaEventRecord(e0, cs);
      for(int q = 0; q < nch; q++)
        cudaMemcpy2DAsync(d + q * csz, sizeof(double) * n, h + q * csz, sizeof(double) * n, sizeof(double) * csz, m, cudaMemcpyHostToDevice, cs);
      cudaEventRecord(e1, cs); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      printf("2-D, %d column chunks: %.1f ms = %.1f GB/s\n", nch, ms, 8e-6 * m * n / ms);
    }
    {
      const int nch = 16; const long long rows = m / nch;
      cudaEventRecord(e0, cs);
      for(int q = 0; q < nch; q++) cudaMemcpyAsync(d + q * rows * n, h + q * rows * n, sizeof(double) * rows * n, cudaMemcpyHostToDevice, cs);
      cudaEventRecord(e1, cs); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      printf("1-D, 16 row chunks: %.1f ms = %.1f GB/s\n", ms, 8e-6 * m * n / ms);
    }
    {
      spin<<<148 * 4, 256, 0, ks>>>(scratch, 1 << 28, 40);
      cudaEventRecord(e0, cs);
      cudaMemcpyAsync(d, h, sizeof(double) * m * n, cudaMemcpyHostToDevice, cs);
      cudaEventRecord(e1, cs); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      cudaDeviceSynchronize();
      printf("contiguous with an HBM-bound kernel running: %.1f ms = %.1f GB/s\n", ms, 8e-6 * m * n / ms);
    }
  }
  // overlap pattern of hb_lowrank_kkt_system_host: per column chunk two 2-D copies + event on the copy stream, wait + kernel on the compute stream
  {
    const int nch = 16; const long long csz = n / nch, half = m / 2;
    cudaEvent_t ev[nch], tk[nch + 1], t0;
    for(int q = 0; q < nch; q++) { cudaEventCreateWithFlags(&ev[q], cudaEventDisableTiming); cudaEventCreate(&tk[q]); }
    cudaEventCreate(&tk[nch]); cudaEventCreate(&t0);
    cudaFuncSetAttribute(spin_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for(int variant = 0; variant < 5; variant++) {
      cudaDeviceSynchronize();
      cudaEventRecord(t0, ks);
      cudaEventRecord(ev[0], ks); cudaStreamWaitEvent(cs, ev[0], 0);
      for(int q = 0; q < nch; q++) {
        if(variant == 3 || variant == 4) { // column chunk as per-row 1-D copies (copy engine), 4 (variant 3) or 8 (variant 4) chunks
          const int nc = variant == 3 ? 4 : 8;
          if(q < nc) {
            const long long cw = n / nc;
            for(long long r = 0; r < m; r++) cudaMemcpyAsync(d + r * n + q * cw, h + r * n + q * cw, sizeof(double) * cw, cudaMemcpyHostToDevice, cs);
          }
        } else if(variant == 0 || variant == 2) {
          cudaMemcpy2DAsync(d + q * csz, sizeof(double) * n, h + q * csz, sizeof(double) * n, sizeof(double) * csz, half, cudaMemcpyHostToDevice, cs);
          cudaMemcpy2DAsync(d + half * n + q * csz, sizeof(double) * n, h + half * n + q * csz, sizeof(double) * n, sizeof(double) * csz, m - half, cudaMemcpyHostToDevice, cs);
        } else { // row chunks, 1-D
          const long long rows = m / nch;
          cudaMemcpyAsync(d + q * rows * n, h + q * rows * n, sizeof(double) * rows * n, cudaMemcpyHostToDevice, cs);
        }
        cudaEventRecord(ev[q], cs);
      }
      for(int q = 0; q < nch; q++) {
        cudaStreamWaitEvent(ks, ev[q], 0);
        cudaEventRecord(tk[q], ks);
        if(variant == 2 || variant == 3 || variant == 4) spin_big<<<148, 384, 200 * 1024, ks>>>(scratch, 1 << 26, 16);
        else spin<<<148 * 4, 256, 0, ks>>>(scratch, 1 << 26, 8);
      }
      cudaEventRecord(tk[nch], ks);
      cudaDeviceSynchronize();
      const char* names[5] = {"2-D column chunks, light kernel", "1-D row chunks, light kernel", "2-D column chunks, 200 KB-smem kernel", "per-row 1-D copies x4 chunks, 200 KB-smem kernel", "per-row 1-D copies x8 chunks, 200 KB-smem kernel"};
      printf("variant %d (%s): chunk ready times:", variant, names[variant]);
      for(int q = 0; q <= nch; q++) { cudaEventElapsedTime(&ms, t0, tk[q]); printf(" %.1f", ms); }
      printf("\n");
    }
  }
  return 0;
}
