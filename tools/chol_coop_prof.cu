// Per-phase cycle counts of the cooperative Cholesky kernel (CTA 0): nvcc -arch=sm_100a -O3 -I../hiop_b200/csrc chol_coop_prof.cu
#include "../hiop_b200/csrc/hb_chol_coop.cu"
#include <vector>
#include <cstdio>
thread_local char g_hb_err[512];
long long g_hb_launches = 0;
int main(int argc, char** argv)
{
  const int N = argc > 1 ? atoi(argv[1]) : 1000;
  std::vector<double> h((size_t)N * N);
  for(int j = 0; j < N; j++)
    for(int i = 0; i < N; i++) h[(size_t)j * N + i] = (i == j ? N : 0.0) + 1.0 / (1.0 + abs(i - j));
  double* A; int* info; long long* prof;
  cudaMalloc(&A, sizeof(double) * h.size()); cudaMalloc(&info, 4); cudaMalloc(&prof, 8 * 10);
  cudaFuncSetAttribute(k_chol_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CoopSmem));
  int lda = N, n = N;
  for(int G : {148, 120}) {
    cudaMemcpy(A, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice);
    cudaMemset(prof, 0, 80); cudaMemset(info, 0, 4);
    double* invd = nullptr;
    void* args[] = {&A, &lda, &n, &info, &invd, &prof};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t rc = cudaLaunchCooperativeKernel((const void*)k_chol_coop, dim3(G), dim3(CT), args, sizeof(CoopSmem), 0);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long p[10]; cudaMemcpy(p, prof, 80, cudaMemcpyDeviceToHost);
    printf("N=%d G=%d rc=%d %.3f ms | cycles: load %lld diag %lld l21 %lld sync1 %lld trail %lld sync2 %lld | diag: chol16 %lld inv %lld below %lld rank16 %lld\n", N, G, (int)rc, ms, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9]);
  }
  return 0;
}
