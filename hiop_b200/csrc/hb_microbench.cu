// In-run peak measurements for bench.py's roofline denominators (the driver's MEASURED_PEAKS.json holds only HBM copy and bf16):
//   * FP64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA) issued back to back from registers by every warp of every SM,
//   * INT8 tcgen05: tcgen05.mma.cta_group::1.kind::i8, M = 128, N = 256, K = 32 issued back to back by one thread per SM on
//     resident shared-memory operands (no loads in the loop): the rate k_oz_gemm is bound by.
// Both are timed with CUDA events on the context stream over a few milliseconds, after a warm-up launch.
#include "hb_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_peak_dmma(double* __restrict__ out, int iters, double a, double b)
{
  double c[8][2];
#pragma unroll
  for(int i = 0; i < 8; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for(int it = 0; it < iters; it++) {
#pragma unroll
    for(int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for(int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one CTA per SM: 16 KB A tile (128 rows x 128 B) + 32 KB B tile (256 rows x 128 B) in the SWIZZLE_128B K-major layout (contents do not
// matter for the rate), two 256-column TMEM accumulators used alternately
__global__ void __launch_bounds__(128, 1)
k_peak_i8(int iters, int* __restrict__ sink)
{
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for(int i = tid; i < (16 + 32) * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 3);
  if(tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s2u(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if(warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s2u(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); // generic-proxy writes of the operands -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if(tid == 0) {
    // K-major SWIZZLE_128B descriptors (hb_ozaki.cu: make_desc_sw); idesc: c = S32, a = b = INT8, N = 256, M = 128
    auto desc = [](uint32_t addr) {
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)1 << 16;
      d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)2 << 61;
      return d;
    };
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t sa = s2u(smem), sb = s2u(smem + 16 * 1024);
    for(int it = 0; it < iters; it++) {
#pragma unroll
      for(int ks = 0; ks < 4; ks++) {
        const uint32_t acc = (it | ks) ? 1u : 0u;
        asm volatile(
            "{\n.reg .pred pp;\nsetp.ne.b32 pp, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, pp;\n}\n" ::"r"(tmem + (uint32_t)((it & 1) * 256)),
            "l"(desc(sa + ks * 32)), "l"(desc(sb + ks * 32)), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(s2u(&bar)) : "memory");
    unsigned ok = 0;
    while(!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s2u(&bar)), "r"(0u) : "memory");
    }
    sink[blockIdx.x] = iters;
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if(warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512));
}

} // namespace

// which: 0 = FP64 DMMA (TFLOP/s), 1 = INT8 tcgen05 (TOP/s, 2 ops per MAC)
extern "C" int hb_microbench_peak(hb_ctx* c, int which, double* result_host)
{
  HB_REQUIRE(c && result_host && (which == 0 || which == 1), "hb_microbench_peak: bad arguments");
  HB_CUDA(cudaSetDevice(c->device));
  cudaEvent_t e0, e1;
  HB_CUDA(cudaEventCreate(&e0));
  HB_CUDA(cudaEventCreate(&e1));
  float ms = 0.f;
  double best = 0.0;
  if(which == 0) {
    const int blocks = c->num_sms * 4, iters = 4000;
    HB_CHECK(hb_ws_reserve(c, sizeof(double) * (size_t)blocks * 256));
    k_peak_dmma<<<blocks, 256, 0, c->stream>>>((double*)c->ws, 100, 1.0000001, 1e-9);
    HB_LAUNCHED();
    for(int rep = 0; rep < 3; rep++) {
      HB_CUDA(cudaEventRecord(e0, c->stream));
      k_peak_dmma<<<blocks, 256, 0, c->stream>>>((double*)c->ws, iters, 1.0000001, 1e-9);
      HB_LAUNCHED();
      HB_CUDA(cudaEventRecord(e1, c->stream));
      HB_CUDA(cudaEventSynchronize(e1));
      HB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double tf = 2.0 * 8 * 256 * (double)iters * 8 * blocks / (ms * 1e-3) / 1e12; // 8 DMMAs of 8x8x4 per warp per iteration, 8 warps
      if(tf > best) best = tf;
    }
  } else {
    const int blocks = c->num_sms, iters = 20000;
    const int smem = 48 * 1024 + 1024;
    static bool attr[16] = {false};
    if(c->device < 16 && !attr[c->device]) {
      HB_CUDA(cudaFuncSetAttribute(k_peak_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr[c->device] = true;
    }
    HB_CHECK(hb_ws_reserve(c, sizeof(int) * (size_t)blocks));
    k_peak_i8<<<blocks, 128, smem, c->stream>>>(200, (int*)c->ws);
    HB_LAUNCHED();
    for(int rep = 0; rep < 3; rep++) {
      HB_CUDA(cudaEventRecord(e0, c->stream));
      k_peak_i8<<<blocks, 128, smem, c->stream>>>(iters, (int*)c->ws);
      HB_LAUNCHED();
      HB_CUDA(cudaEventRecord(e1, c->stream));
      HB_CUDA(cudaEventSynchronize(e1));
      HB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double tops = 2.0 * 128 * 256 * 32 * 4.0 * (double)iters * blocks / (ms * 1e-3) / 1e12;
      if(tops > best) best = tops;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *result_host = best;
  return HB_OK;
}
