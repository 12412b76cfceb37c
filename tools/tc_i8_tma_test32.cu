// Stage-2 probe: TMA (3-D tensor map, SWIZZLE_64B) -> shared -> tcgen05.mma.kind::i8 (M=128, N=64, K=32B) -> TMEM -> tcgen05.ld,
// with the producer / MMA / epilogue warp roles and full/empty mbarriers of the production kernel.
// Computes C(128x64) = sum_{(p,q) in pairs} Q_p[rowsA] * Q_q[rowsB]^T into separate TMEM accumulators per t = p+q and checks each.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_i8_tma_test tools/tc_i8_tma_test.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while(0)

constexpr int S = 3;                 // slices
constexpr int MROWS = 256;           // rows of the sliced matrix
constexpr int KBYTES = 512;          // K extent in bytes (int8)
constexpr int KS = 32;
constexpr int STAGES = 4;
constexpr int TM = 128, TN = 64;
constexpr int A_TILE = TM * KS, B_TILE = TN * KS;          // per slice
constexpr int STAGE_BYTES = S * (A_TILE + B_TILE);

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s2u(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s2u(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(s2u(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity)
{
  asm volatile(
      "{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni WD;\nbra.uni WL;\nWD:\n}\n" ::"r"(s2u(b)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(s2u(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(s2u(bar))
               : "memory");
}
// K-major SWIZZLE_64B operand descriptor: rows at 64 B, 8-row groups at SBO = 512 B
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr)
{
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major) = 1
  d |= (uint64_t)((256 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // version
  d |= (uint64_t)6 << 61;
  return d;
}

__global__ void __launch_bounds__(256) k_test(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int rowA0, int rowB0,
                                              int32_t* __restrict__ C /* [S][128][64] per t */)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage_base = smem;                                   // STAGES * STAGE_BYTES
  unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + STAGES * STAGE_BYTES);
  unsigned long long* empty = full + STAGES;
  unsigned long long* accfull = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if(tid == 0) {
    for(int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(accfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if(warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s2u(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int nk = KBYTES / KS;

  if(warp == 0 && lane == 0) {
    // ---------------- TMA producer ----------------
    for(int it = 0; it < nk; it++) {
      const int s = it % STAGES;
      const unsigned ph = (it / STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_expect_tx(&full[s], STAGE_BYTES);
      uint8_t* st = stage_base + s * STAGE_BYTES;
      for(int p = 0; p < S; p++) {
        tma_load_3d(st + p * A_TILE, &mapA, it * KS, rowA0, p, &full[s]);
        tma_load_3d(st + S * A_TILE + p * B_TILE, &mapB, it * KS, rowB0, p, &full[s]);
      }
    }
  } else if(warp == 1 && lane == 0) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    for(int it = 0; it < nk; it++) {
      const int s = it % STAGES;
      const unsigned ph = (it / STAGES) & 1;
      mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t sa = s2u(stage_base + s * STAGE_BYTES), sb = sa + S * A_TILE;
      for(int p = 0; p < S; p++)
        for(int q = 0; p + q < S; q++) {
          const int t = p + q;
          for(int ks = 0; ks < KS / 32; ks++) {
            const uint64_t da = make_desc_sw64(sa + p * A_TILE + ks * 32);
            const uint64_t db = make_desc_sw64(sb + q * B_TILE + ks * 32);
            const uint32_t acc = (it > 0 || ks > 0 || p > 0) ? 1u : 0u;   // first MMA into accumulator t: (it=0, ks=0, p=0)
            asm volatile(
                "{\n.reg .pred pp;\nsetp.ne.b32 pp, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, pp;\n}\n" ::"r"(tmem + (uint32_t)(t * TN)),
                "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
                : "memory");
          }
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s2u(&empty[s])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s2u(accfull)) : "memory");
  } else if(warp >= 4) {
    // ---------------- epilogue warps: TMEM lanes 32*(warp%4) ----------------
    mbar_wait(accfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int wq = warp & 3;
    for(int t = 0; t < S; t++)
      for(int c0 = 0; c0 < TN; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)(t * TN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        const int row = wq * 32 + lane;
        for(int j = 0; j < 32; j++) C[(t * TM + row) * TN + c0 + j] = (int32_t)v[j];
      }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if(warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
  std::vector<int8_t> Q((size_t)S * MROWS * KBYTES);
  srand(3);
  for(auto& x : Q) x = (int8_t)(rand() % 255 - 127);
  int8_t* dQ;
  int32_t* dC;
  CK(cudaMalloc(&dQ, Q.size()));
  CK(cudaMemcpy(dQ, Q.data(), Q.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dC, sizeof(int32_t) * S * TM * TN));
  CK(cudaMemset(dC, 0xff, sizeof(int32_t) * S * TM * TN));
  PFN_encodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if(!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap mapA, mapB;
  cuuint64_t dims[3] = {(cuuint64_t)KBYTES, (cuuint64_t)MROWS, (cuuint64_t)S};
  cuuint64_t strides[2] = {(cuuint64_t)KBYTES, (cuuint64_t)KBYTES * MROWS};
  cuuint32_t boxA[3] = {KS, TM, 1}, boxB[3] = {KS, TN, 1}, es[3] = {1, 1, 1};
  CUresult r1 = encode(&mapA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dQ, dims, strides, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = encode(&mapB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, dQ, dims, strides, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if(r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
  const int rowA0 = 128, rowB0 = 64;
  const size_t smem = STAGES * STAGE_BYTES + 256;
  CK(cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_test<<<1, 256, smem>>>(mapA, mapB, rowA0, rowB0, dC);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> C((size_t)S * TM * TN);
  CK(cudaMemcpy(C.data(), dC, sizeof(int32_t) * C.size(), cudaMemcpyDeviceToHost));
  long bad = 0;
  for(int t = 0; t < S; t++)
    for(int i = 0; i < TM; i++)
      for(int j = 0; j < TN; j++) {
        long s = 0;
        for(int p = 0; p <= t; p++) {
          const int q = t - p;
          const int8_t* a = &Q[((size_t)p * MROWS + rowA0 + i) * KBYTES];
          const int8_t* b = &Q[((size_t)q * MROWS + rowB0 + j) * KBYTES];
          for(int k = 0; k < KBYTES; k++) s += (int)a[k] * (int)b[k];
        }
        if((int32_t)s != C[(t * TM + i) * TN + j]) {
          if(bad < 6) printf("mismatch t=%d (%d,%d): got %d want %ld\n", t, i, j, C[(t * TM + i) * TN + j], s);
          bad++;
        }
      }
  printf("tcgen05 i8 + TMA SWIZZLE_64B probe: %ld mismatches of %d\n", bad, S * TM * TN);
  return bad ? 2 : 0;
}
