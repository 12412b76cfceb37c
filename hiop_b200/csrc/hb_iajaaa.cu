// write_kkt interchange files: the .iajaaa text format HiOp uses to dump KKT systems
// (src/Utils/hiopCSR_IO.hpp:89-155 writer for dense matrices, src/LinAlg/csr_iajaaa.md format, src/LinAlg/load_kkt_mat.m reader).
// Byte-compatible with the reference writer: same zero threshold (1e-25 on the upper triangle), 1-based indices, "%.20f " values,
// rhs / solution vectors appended as one line each.
#include "hb_common.cuh"
#include <cstdio>
#include <cmath>
#include <vector>

extern "C" int hb_iajaaa_write_matrix_host(const char* filename, int N, const double* M, int nx, int meq, int mineq)
{
  HB_REQUIRE(filename && N >= 0 && (N == 0 || M), "hb_iajaaa_write_matrix_host: bad arguments");
  FILE* f = fopen(filename, "w+");
  if(!f) return hb_fail(HB_ERR_INVALID, "hb_iajaaa_write_matrix_host: cannot open '%s'", filename);
  const double zero_tol = 1e-25; // hiopCSR_IO.hpp:111
  int nnz = 0;
  for(int i = 0; i < N; i++)
    for(int j = i; j < N; j++)
      if(std::fabs(M[(size_t)i * N + j]) > zero_tol) nnz++;
  fprintf(f, "%d\n%d\n%d\n%d\n%d\n", N, nx, meq, mineq, nnz);
  int offset = 1;
  fprintf(f, "%d ", offset);
  for(int i = 0; i < N; i++) {
    for(int j = i; j < N; j++)
      if(std::fabs(M[(size_t)i * N + j]) > zero_tol) offset++;
    fprintf(f, "%d ", offset);
  }
  fprintf(f, "\n");
  for(int i = 0; i < N; i++)
    for(int j = i; j < N; j++)
      if(std::fabs(M[(size_t)i * N + j]) > zero_tol) fprintf(f, "%d ", j + 1);
  fprintf(f, "\n");
  for(int i = 0; i < N; i++)
    for(int j = i; j < N; j++)
      if(std::fabs(M[(size_t)i * N + j]) > zero_tol) fprintf(f, "%.20f ", M[(size_t)i * N + j]);
  fprintf(f, "\n");
  fclose(f);
  return HB_OK;
}

extern "C" int hb_iajaaa_append_vector_host(const char* filename, int N, const double* v)
{
  HB_REQUIRE(filename && N >= 0 && (N == 0 || v), "hb_iajaaa_append_vector_host: bad arguments");
  FILE* f = fopen(filename, "a+"); // hiopCSR_IO.hpp:55
  if(!f) return hb_fail(HB_ERR_INVALID, "hb_iajaaa_append_vector_host: cannot open '%s'", filename);
  for(int i = 0; i < N; i++) fprintf(f, "%.20f ", v[i]);
  fprintf(f, "\n");
  fclose(f);
  return HB_OK;
}

extern "C" int hb_iajaaa_write_matrix(hb_ctx* c, const char* filename, int N, const double* M_dev, int nx, int meq, int mineq)
{
  HB_REQUIRE(c && (N == 0 || M_dev), "hb_iajaaa_write_matrix: bad arguments");
  std::vector<double> h((size_t)N * N);
  if(N) {
    HB_CUDA(cudaMemcpyAsync(h.data(), M_dev, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return hb_iajaaa_write_matrix_host(filename, N, h.data(), nx, meq, mineq);
}

extern "C" int hb_iajaaa_append_vector(hb_ctx* c, const char* filename, int N, const double* v_dev)
{
  HB_REQUIRE(c && (N == 0 || v_dev), "hb_iajaaa_append_vector: bad arguments");
  std::vector<double> h((size_t)N);
  if(N) {
    HB_CUDA(cudaMemcpyAsync(h.data(), v_dev, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return hb_iajaaa_append_vector_host(filename, N, h.data());
}
