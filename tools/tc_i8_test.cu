// Stage-1 probe for the INT8-slice (Ozaki) condensation: one CTA computes C(128x128,int32) = A(128xK,int8) * B(128xK,int8)^T with
// tcgen05.mma.kind::i8, accumulator in TMEM, operands in the canonical no-swizzle K-major shared-memory layout
// (core matrix = 8 rows x 16 bytes), and checks it against the CPU. Validates descriptors / TMEM plumbing.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_i8_test tools/tc_i8_test.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while(0)

constexpr int M = 128, N = 128, KB = 128;          // KB bytes of K
constexpr int KSTEP = 32;                          // bytes of K per tcgen05.mma.kind::i8
constexpr int TILE_BYTES = 128 * KSTEP;            // one K-step tile of 128 rows
// canonical K-major, no swizzle: element (r, kbyte) of a K-step tile at (r/8)*SBO + (kbyte/16)*LBO + (r%8)*16 + kbyte%16
constexpr int LBO = 128, SBO = 256;

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr)
{
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);            // start address, bits [0,14)
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;            // leading byte offset, bits [16,30)
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;            // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                                // version = 1 (Blackwell)
  // base_offset = 0, lbo_mode = 0, layout_type (bits 61-63) = 0 (SWIZZLE_NONE)
  return d;
}

__global__ void __launch_bounds__(128) k_test(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ C)
{
  __shared__ __align__(1024) uint8_t sA[KB / KSTEP][TILE_BYTES];
  __shared__ __align__(1024) uint8_t sB[KB / KSTEP][TILE_BYTES];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for(int e = tid; e < M * KB; e += 128) {
    const int r = e / KB, kb = e % KB;
    const int ks = kb / KSTEP, kk = kb % KSTEP;
    const int off = (r / 8) * SBO + (kk / 16) * LBO + (r % 8) * 16 + kk % 16;
    sA[ks][off] = (uint8_t)A[r * KB + kb];
    sB[ks][off] = (uint8_t)B[r * KB + kb];
  }
  if(tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if(warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy smem writes -> visible to the async (tensor core) proxy
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base;

  if(tid == 0) {
    // instruction descriptor: c_format=S32 (2) bits[4,6); a_format=INT8 (1) bits[7,10); b_format=INT8 (1) bits[10,13); K-major both;
    // n_dim = N>>3 bits[17,23); m_dim = M>>4 bits[24,29)
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for(int ks = 0; ks < KB / KSTEP; ks++) {
      const uint64_t da = make_desc((uint32_t)__cvta_generic_to_shared(&sA[ks][0]));
      const uint64_t db = make_desc((uint32_t)__cvta_generic_to_shared(&sB[ks][0]));
      const uint32_t acc = ks > 0 ? 1u : 0u;
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "setp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
          "}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W1:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra.uni D1;\n"
        "bra.uni W1;\n"
        "D1:\n"
        "}\n" ::"r"(addr) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // warp w reads TMEM lanes 32w..32w+31 (= rows), 32 columns at a time
  for(int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    const int row = warp * 32 + lane;
    for(int j = 0; j < 32; j++) C[row * N + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if(warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128));
}

int main()
{
  std::vector<int8_t> A(M * KB), B(N * KB);
  srand(1);
  for(auto& x : A) x = (int8_t)(rand() % 128 - 64);
  for(auto& x : B) x = (int8_t)(rand() % 128 - 64);
  int8_t *dA, *dB;
  int32_t* dC;
  CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dC, sizeof(int32_t) * M * N));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, sizeof(int32_t) * M * N));
  k_test<<<1, 128>>>(dA, dB, dC);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> C(M * N);
  CK(cudaMemcpy(C.data(), dC, sizeof(int32_t) * M * N, cudaMemcpyDeviceToHost));
  long bad = 0;
  for(int i = 0; i < M; i++)
    for(int j = 0; j < N; j++) {
      int32_t s = 0;
      for(int k = 0; k < KB; k++) s += (int32_t)A[i * KB + k] * (int32_t)B[j * KB + k];
      if(s != C[i * N + j]) {
        if(bad < 8) printf("mismatch (%d,%d): got %d want %d\n", i, j, C[i * N + j], s);
        bad++;
      }
    }
  printf("tcgen05 kind::i8 probe: %ld mismatches of %d\n", bad, M * N);
  return bad ? 2 : 0;
}
