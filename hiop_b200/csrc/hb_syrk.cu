// C = A * diag(d) * A^T in FP64 on the sm_100a DMMA pipe -- the condensation kernel.
//
// Replaces the reference's scalar triple loops symmMatTimesDiagTimesMatTrans_local / matTimesDiagTimesMatTrans_local
// (src/Optimization/hiopHessianLowRank.cpp:1079-1154) which stream J m/2 times from DRAM. Here the augmented row set
// A = [J; S_t; Y_t] (M = m + 2l rows, K = n_local columns, K-contiguous rows living in several buffers -> a device
// table of row pointers) is read ONCE per output tile pair and all of W = J D J^T, S1, Y1 and the three l x l
// blocks of V come out of the same pass.
//
// Why DMMA and not tcgen05: tcgen05.mma has no f64 kind; mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) is the FP64
// tensor path of sm_100a and measures 37.1 TFLOP/s on B200 (tools/microbench_fp64.cu) = 64 FMA/clk/SM at 1965 MHz.
//
// Work decomposition: 128x128 output tiles (upper triangle of the tile grid only), K swept in BK=16 chunks through a
// 4-stage cp.async pipeline into padded shared memory (row stride 20 doubles = 160 B -> the m8n8k4 fragment reads
// of a half-warp, 4 rows x 32 B, fall into 4 distinct 32 B bank groups: conflict-free LDS.64).
// The (tile, K-range) space is cut into one contiguous range per CTA by a host-side schedule (stream-K): every SM
// gets the same number of MMA iterations whatever M is. Each CTA writes its partial 128x128 tile to a workspace
// slot; a second kernel sums the slots of each tile in a FIXED order and mirrors the result, so the output is
// bit-reproducible run to run (no atomics).
#include "hb_common.cuh"
#include <cstdlib>

namespace {

constexpr int BM = 128;          // tile rows = tile cols
constexpr int BK = 16;           // doubles per K chunk
constexpr int STAGES = 4;
constexpr int THREADS = 256;     // 8 warps: 2 (rows) x 4 (cols), warp tile 64 x 32
constexpr int LDS_ROW = BK + 4;  // padded row stride in doubles (160 B)
constexpr int TILE_D = BM * LDS_ROW;

struct Stage
{
  double a[TILE_D];
  double b[TILE_D];
  double d[BK];
};
constexpr size_t SMEM_BYTES = sizeof(Stage) * STAGES + 2 * BM * sizeof(const double*);

struct Seg
{
  int ti, tj;      // tile coordinates, ti <= tj
  int k_begin;     // first K iteration (units of BK columns)
  int k_count;     // number of K iterations
  int slot;        // workspace slot receiving the partial tile
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes)
{
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes)
{
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Loads one K chunk (BK columns starting at column k0) of the row tile(s) into a stage.
template <bool ALIGN16>
__device__ __forceinline__ void load_stage(Stage& st, const double* const* srow_a, const double* const* srow_b, bool diag,
                                           const double* __restrict__ dvec_or_null, const double* __restrict__ dummy, long long k0, long long K)
{
  const int tid = threadIdx.x;
  const double* dvec = dummy; // only used as a valid address for zero-byte copies
  if(ALIGN16) {
    const int kc = tid & 7;            // 16-byte chunk within the row
    const long long k = k0 + kc * 2;
    long long rem = K - k;
    const int nb = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
    const long long koff = nb ? k : 0; // keep the address valid when nothing is read
#pragma unroll
    for(int j = 0; j < 4; j++) {
      const int row = (tid >> 3) + 32 * j;
      const double* pa = srow_a[row];
      cp_async16(&st.a[row * LDS_ROW + kc * 2], pa ? pa + koff : (const double*)dvec, pa ? nb : 0);
      if(!diag) {
        const double* pb = srow_b[row];
        cp_async16(&st.b[row * LDS_ROW + kc * 2], pb ? pb + koff : (const double*)dvec, pb ? nb : 0);
      }
    }
    if(tid < 8) {
      if(dvec_or_null) {
        const long long kd = k0 + tid * 2;
        long long r2 = K - kd;
        const int nbd = r2 >= 2 ? 16 : (r2 == 1 ? 8 : 0);
        cp_async16(&st.d[tid * 2], nbd ? dvec_or_null + kd : dummy, nbd);
      } else {
        st.d[tid * 2] = 1.0;
        st.d[tid * 2 + 1] = 1.0;
      }
    }
  } else {
    const int kc = tid & 15;           // 8-byte chunk within the row
    const long long k = k0 + kc;
    const int nb = k < K ? 8 : 0;
    const long long koff = nb ? k : 0;
#pragma unroll
    for(int j = 0; j < 8; j++) {
      const int row = (tid >> 4) + 16 * j;
      const double* pa = srow_a[row];
      cp_async8(&st.a[row * LDS_ROW + kc], pa ? pa + koff : (const double*)dvec, pa ? nb : 0);
      if(!diag) {
        const double* pb = srow_b[row];
        cp_async8(&st.b[row * LDS_ROW + kc], pb ? pb + koff : (const double*)dvec, pb ? nb : 0);
      }
    }
    if(tid < 16) {
      if(dvec_or_null) {
        const long long kd = k0 + tid;
        const int nbd = kd < K ? 8 : 0;
        cp_async8(&st.d[tid], nbd ? dvec_or_null + kd : dummy, nbd);
      } else {
        st.d[tid] = 1.0;
      }
    }
  }
}

template <bool ALIGN16>
__global__ void __launch_bounds__(THREADS, 1)
k_syrk_diag(const double* const* __restrict__ rowptr, int M, long long K, const double* __restrict__ dvec, const Seg* __restrict__ segs,
            const int* __restrict__ cta_seg_begin, double* __restrict__ ws)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Stage* stages = reinterpret_cast<Stage*>(smem_raw);
  const double** srow_a = reinterpret_cast<const double**>(smem_raw + sizeof(Stage) * STAGES);
  const double** srow_b = srow_a + BM;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int warp_m = warp & 1, warp_n = warp >> 1;
  const int g = lane >> 2, t4 = lane & 3;

  const int sb = cta_seg_begin[blockIdx.x], se = cta_seg_begin[blockIdx.x + 1];
  for(int si = sb; si < se; si++) {
    const Seg sg = segs[si];
    const bool diag = sg.ti == sg.tj;
    __syncthreads(); // previous segment fully consumed before the row tables / stages are reused
    if(tid < BM) {
      const int ra = sg.ti * BM + tid;
      srow_a[tid] = ra < M ? rowptr[ra] : nullptr;
    } else {
      const int rb = sg.tj * BM + (tid - BM);
      srow_b[tid - BM] = rb < M ? rowptr[rb] : nullptr;
    }
    __syncthreads();

    double acc[8][4][2];
#pragma unroll
    for(int i = 0; i < 8; i++)
#pragma unroll
      for(int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int kcount = sg.k_count;
    const long long kbase = (long long)sg.k_begin * BK;
#pragma unroll
    for(int s = 0; s < STAGES - 1; s++) {
      if(s < kcount) load_stage<ALIGN16>(stages[s], srow_a, srow_b, diag, dvec, (const double*)rowptr, kbase + (long long)s * BK, K);
      cp_async_commit();
    }
    for(int it = 0; it < kcount; it++) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      {
        const int nx = it + STAGES - 1;
        if(nx < kcount) load_stage<ALIGN16>(stages[nx % STAGES], srow_a, srow_b, diag, dvec, (const double*)rowptr, kbase + (long long)nx * BK, K);
        cp_async_commit();
      }
      const Stage& st = stages[it % STAGES];
      const double* sA = st.a + (warp_m * 64 + g) * LDS_ROW + t4;
      const double* sB = (diag ? st.a : st.b) + (warp_n * 32 + g) * LDS_ROW + t4;
#pragma unroll
      for(int kk = 0; kk < BK / 4; kk++) {
        const double dv = st.d[kk * 4 + t4];
        double af[8], bf[4];
#pragma unroll
        for(int i = 0; i < 8; i++) af[i] = sA[i * 8 * LDS_ROW + kk * 4];
#pragma unroll
        for(int j = 0; j < 4; j++) bf[j] = sB[j * 8 * LDS_ROW + kk * 4] * dv;
#pragma unroll
        for(int i = 0; i < 8; i++)
#pragma unroll
          for(int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    }
    cp_async_wait<0>();

    double* slot = ws + (size_t)sg.slot * (BM * BM);
#pragma unroll
    for(int i = 0; i < 8; i++) {
      const int row = warp_m * 64 + i * 8 + g;
#pragma unroll
      for(int j = 0; j < 4; j++) {
        const int col = warp_n * 32 + j * 8 + t4 * 2;
        *reinterpret_cast<double2*>(slot + row * BM + col) = make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Warp-specialised variant (the fast path): 8 MMA warps + 4 producer warps, mbarrier full/empty ring.
// The producers issue 16-byte cp.async copies (zero-filling rows beyond M and the K tail through the src-size operand)
// and signal the stage's mbarrier with cp.async.mbarrier.arrive.noinc; the MMA warps never execute a CTA-wide barrier
// and never compute a global address, so they drift out of phase and keep the FP64 tensor pipe busy during refills.
// (A first version staged rows with 256-byte cp.async.bulk copies from one producer warp: 288 bulk copies per stage
// made the producer the bottleneck -- 61 ms vs 40 ms -- see profiles/README_r01.md.)
// Needs 16-byte aligned rows (else k_syrk_diag<false> runs).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int WBK = 32;                 // doubles per K chunk
constexpr int WSTAGES = 3;
constexpr int WLDS = WBK + 4;           // 36 doubles = 288 B row stride (288 mod 128 = 32 -> conflict-free fragment reads)
constexpr int WTILE_D = BM * WLDS;
constexpr int WPROD = 128;              // producer threads (4 warps)
constexpr int WTHREADS = 256 + WPROD;
struct WStage
{
  double a[WTILE_D];
  double b[WTILE_D];
  double d[WBK];
};
constexpr size_t WSMEM_BYTES = sizeof(WStage) * WSTAGES + 2 * BM * sizeof(const double*) + 2 * WSTAGES * sizeof(unsigned long long);

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_cp_async(unsigned long long* bar)
{
  // arrives (without incrementing the pending count) once all prior cp.async of this thread have landed
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(addr), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(WTHREADS, 1)
k_syrk_ws(const double* const* __restrict__ rowptr, int M, long long K, const double* __restrict__ dvec, const Seg* __restrict__ segs,
          const int* __restrict__ cta_seg_begin, double* __restrict__ ws)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WStage* stages = reinterpret_cast<WStage*>(smem_raw);
  const double** srow = reinterpret_cast<const double**>(smem_raw + sizeof(WStage) * WSTAGES); // [2*BM] row pointers (producers only)
  unsigned long long* full = reinterpret_cast<unsigned long long*>(srow + 2 * BM);
  unsigned long long* empty = full + WSTAGES;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  if(tid == 0) {
#pragma unroll
    for(int s = 0; s < WSTAGES; s++) {
      mbar_init(&full[s], WPROD); // every producer thread arrives once its copies of the stage have landed
      mbar_init(&empty[s], 8);    // one arrival per MMA warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  const int sb = cta_seg_begin[blockIdx.x], se = cta_seg_begin[blockIdx.x + 1];
  int stage = 0;
  unsigned phase = 0;

  if(warp >= 8) {
    // ================= producer warps =================
    const int p = tid - 256;          // 0..127
    const int kc = p & 15;            // 16-byte chunk within the 256-byte row segment
    const int r0 = p >> 4;            // rows r0 + 8*j
    for(int si = sb; si < se; si++) {
      const Seg sg = segs[si];
      const bool diag = sg.ti == sg.tj;
      asm volatile("bar.sync 1, %0;\n" ::"n"(WPROD) : "memory"); // all producers done with the previous segment's row table
      for(int r = p; r < 2 * BM; r += WPROD) {
        const int grow = (r < BM ? sg.ti * BM + r : sg.tj * BM + (r - BM));
        srow[r] = grow < M ? rowptr[grow] : nullptr;
      }
      asm volatile("bar.sync 1, %0;\n" ::"n"(WPROD) : "memory");
      for(int it = 0; it < sg.k_count; it++) {
        const long long k = ((long long)sg.k_begin + it) * WBK + kc * 2;
        const long long rem = K - k;
        const int nb = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
        const long long koff = nb ? k : 0;
        mbar_wait(&empty[stage], phase ^ 1);
        WStage& st = stages[stage];
#pragma unroll 4
        for(int j = 0; j < 16; j++) {
          const int row = r0 + 8 * j;
          const double* pa = srow[row];
          cp_async16(&st.a[row * WLDS + kc * 2], pa ? pa + koff : (const double*)rowptr, pa ? nb : 0);
          if(!diag) {
            const double* pb = srow[BM + row];
            cp_async16(&st.b[row * WLDS + kc * 2], pb ? pb + koff : (const double*)rowptr, pb ? nb : 0);
          }
        }
        if(p < 16 && dvec) cp_async16(&st.d[kc * 2], nb ? dvec + k : (const double*)rowptr, nb);
        mbar_arrive_on_cp_async(&full[stage]);
        if(++stage == WSTAGES) { stage = 0; phase ^= 1; }
      }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  } else {
    // ================= 8 MMA warps =================
    const int warp_m = warp & 1, warp_n = warp >> 1;
    const int g = lane >> 2, t4 = lane & 3;
    const bool unit_d = (dvec == nullptr);
    for(int si = sb; si < se; si++) {
      const Seg sg = segs[si];
      const bool diag = sg.ti == sg.tj;
      double acc[8][4][2];
#pragma unroll
      for(int i = 0; i < 8; i++)
#pragma unroll
        for(int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
      for(int it = 0; it < sg.k_count; it++) {
        mbar_wait(&full[stage], phase);
        const WStage& st = stages[stage];
        const double* sA = st.a + (warp_m * 64 + g) * WLDS + t4;
        const double* sB = (diag ? st.a : st.b) + (warp_n * 32 + g) * WLDS + t4;
#pragma unroll
        for(int kk = 0; kk < WBK / 4; kk++) {
          const double dv = unit_d ? 1.0 : st.d[kk * 4 + t4];
          double af[8], bf[4];
#pragma unroll
          for(int i = 0; i < 8; i++) af[i] = sA[i * 8 * WLDS + kk * 4];
#pragma unroll
          for(int j = 0; j < 4; j++) bf[j] = sB[j * 8 * WLDS + kk * 4] * dv;
#pragma unroll
          for(int i = 0; i < 8; i++)
#pragma unroll
            for(int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        __syncwarp();
        if(lane == 0) mbar_arrive(&empty[stage]);
        if(++stage == WSTAGES) { stage = 0; phase ^= 1; }
      }
      double* slot = ws + (size_t)sg.slot * (BM * BM);
#pragma unroll
      for(int i = 0; i < 8; i++) {
        const int row = warp_m * 64 + i * 8 + g;
#pragma unroll
        for(int j = 0; j < 4; j++) {
          const int col = warp_n * 32 + j * 8 + t4 * 2;
          *reinterpret_cast<double2*>(slot + row * BM + col) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
      }
    }
  }
}

// Sums the partial slots of each tile in schedule order and writes C (both triangles).
__global__ void __launch_bounds__(256)
k_syrk_fixup(int M, const int2* __restrict__ tile_ij, const int* __restrict__ tile_slot_begin, const int* __restrict__ tile_slots,
             const double* __restrict__ ws, double* __restrict__ C, int ldc)
{
  const int t = blockIdx.x;
  const int2 ij = tile_ij[t];
  const int s0 = tile_slot_begin[t], s1 = tile_slot_begin[t + 1];
  const int r0 = blockIdx.y * 16; // 8 row-chunks of 16 rows
  for(int e = threadIdx.x; e < 16 * BM; e += 256) {
    const int r = r0 + e / BM, c = e % BM;
    const int gi = ij.x * BM + r, gj = ij.y * BM + c;
    if(gi >= M || gj >= M) continue;
    if(ij.x == ij.y && c < r) continue; // diagonal tile: use the upper part and mirror it (exact symmetry)
    double v = 0.0;
    for(int s = s0; s < s1; s++) v += ws[(size_t)tile_slots[s] * (BM * BM) + r * BM + c];
    C[(size_t)gi * ldc + gj] = v;
    C[(size_t)gj * ldc + gi] = v;
  }
}

struct Schedule
{
  int M = -1;
  long long K = -1;
  int bk = 0;      // K-chunk (columns per iteration) the schedule counts in
  int G = 0;       // SM count the schedule was built for
  int Gl = 0;      // CTAs to launch
  long long stamp = 0;
  int ntiles = 0, nslots = 0;
  int *d_cta_seg_begin = nullptr, *d_tile_slot_begin = nullptr, *d_tile_slots = nullptr;
  int2* d_tile_ij = nullptr;
  Seg* d_segs = nullptr;
};
constexpr int SCHED_WAYS = 4;
Schedule g_sched[16][SCHED_WAYS]; // per device, small LRU cache keyed by (M, K)
long long g_sched_clock = 0;

int build_schedule(hb_ctx* c, Schedule*& Sout, int M, long long K, int bk)
{
  Schedule* ways = g_sched[c->device];
  int victim = 0;
  for(int w = 0; w < SCHED_WAYS; w++) {
    if(ways[w].M == M && ways[w].K == K && ways[w].bk == bk && ways[w].G == c->num_sms) {
      ways[w].stamp = ++g_sched_clock;
      Sout = &ways[w];
      return HB_OK;
    }
    if(ways[w].stamp < ways[victim].stamp) victim = w;
  }
  Schedule& S = ways[victim];
  Sout = &S;
  S.stamp = ++g_sched_clock;
  HB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(S.d_cta_seg_begin); cudaFree(S.d_tile_slot_begin); cudaFree(S.d_tile_slots); cudaFree(S.d_tile_ij); cudaFree(S.d_segs);
  const int T = (M + BM - 1) / BM;
  const int ntiles = T * (T + 1) / 2;
  const long long kiters = (K + bk - 1) / bk;
  const long long total = (long long)ntiles * kiters;
  int G = c->num_sms;
  if(total < G) G = (int)(total > 0 ? total : 1);
  std::vector<int2> tij(ntiles);
  {
    int t = 0;
    for(int i = 0; i < T; i++)
      for(int j = i; j < T; j++) tij[t++] = make_int2(i, j);
  }
  std::vector<Seg> segs;
  std::vector<int> cta_begin(G + 1, 0);
  std::vector<std::vector<int>> per_tile(ntiles);
  for(int cta = 0; cta < G; cta++) {
    cta_begin[cta] = (int)segs.size();
    long long it = hb_part_begin(total, G, cta), end = hb_part_begin(total, G, cta + 1);
    while(it < end) {
      const int tile = (int)(it / kiters);
      const long long kk0 = it % kiters;
      long long cnt = kiters - kk0;
      if(cnt > end - it) cnt = end - it;
      Seg s;
      s.ti = tij[tile].x; s.tj = tij[tile].y; s.k_begin = (int)kk0; s.k_count = (int)cnt; s.slot = (int)segs.size();
      per_tile[tile].push_back(s.slot);
      segs.push_back(s);
      it += cnt;
    }
  }
  cta_begin[G] = (int)segs.size();
  std::vector<int> tsb(ntiles + 1, 0), tsl;
  for(int t = 0; t < ntiles; t++) {
    tsb[t] = (int)tsl.size();
    for(int s : per_tile[t]) tsl.push_back(s);
  }
  tsb[ntiles] = (int)tsl.size();
  if(tsl.empty()) tsl.push_back(0);
  if(segs.empty()) segs.push_back(Seg{0, 0, 0, 0, 0});
  HB_CUDA(cudaMalloc(&S.d_cta_seg_begin, sizeof(int) * (G + 1)));
  HB_CUDA(cudaMalloc(&S.d_tile_slot_begin, sizeof(int) * (ntiles + 1)));
  HB_CUDA(cudaMalloc(&S.d_tile_slots, sizeof(int) * tsl.size()));
  HB_CUDA(cudaMalloc(&S.d_tile_ij, sizeof(int2) * ntiles));
  HB_CUDA(cudaMalloc(&S.d_segs, sizeof(Seg) * segs.size()));
  HB_CUDA(cudaMemcpy(S.d_cta_seg_begin, cta_begin.data(), sizeof(int) * (G + 1), cudaMemcpyHostToDevice));
  HB_CUDA(cudaMemcpy(S.d_tile_slot_begin, tsb.data(), sizeof(int) * (ntiles + 1), cudaMemcpyHostToDevice));
  HB_CUDA(cudaMemcpy(S.d_tile_slots, tsl.data(), sizeof(int) * tsl.size(), cudaMemcpyHostToDevice));
  HB_CUDA(cudaMemcpy(S.d_tile_ij, tij.data(), sizeof(int2) * ntiles, cudaMemcpyHostToDevice));
  HB_CUDA(cudaMemcpy(S.d_segs, segs.data(), sizeof(Seg) * segs.size(), cudaMemcpyHostToDevice));
  S.M = M; S.K = K; S.bk = bk; S.G = c->num_sms; S.Gl = G; S.ntiles = ntiles; S.nslots = (int)cta_begin[G];
  return HB_OK;
}

bool g_attr_set = false;

} // namespace

// rowptr: DEVICE table of M row pointers (each row K doubles, K-contiguous). d: length K (device) or NULL (= ones).
// C: M x M (ldc), both triangles written. `aligned16`: every row pointer and d are 16-byte aligned.
int hb_syrk_rows(hb_ctx* c, int M, long long K, const double* const* rowptr_dev, bool aligned16, const double* d, double* C, int ldc)
{
  HB_REQUIRE(c && M >= 0 && K >= 0 && ldc >= M, "hb_syrk_rows: bad arguments");
  if(M == 0) return HB_OK;
  if(K == 0) {
    HB_CUDA(cudaMemset2DAsync(C, sizeof(double) * ldc, 0, sizeof(double) * M, M, c->stream));
    return HB_OK;
  }
  HB_REQUIRE(c->device < 16, "device ordinal too large");
  const bool d_aligned = (reinterpret_cast<uintptr_t>(d) & 15u) == 0;
  const char* force_generic = getenv("HB_SYRK_GENERIC");
  const bool use_ws = aligned16 && d_aligned && ((reinterpret_cast<uintptr_t>(rowptr_dev) & 15u) == 0) && !(force_generic && force_generic[0] == '1');
  Schedule* Sp = nullptr;
  HB_CHECK(build_schedule(c, Sp, M, K, use_ws ? WBK : BK));
  Schedule& S = *Sp;
  HB_CHECK(hb_ws_reserve(c, (size_t)S.nslots * BM * BM * sizeof(double)));
  if(!g_attr_set) {
    HB_CUDA(cudaFuncSetAttribute(k_syrk_diag<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    HB_CUDA(cudaFuncSetAttribute(k_syrk_diag<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    HB_CUDA(cudaFuncSetAttribute(k_syrk_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WSMEM_BYTES));
    g_attr_set = true;
  }
  const int G = S.Gl;
  if(c->timing) HB_CUDA(cudaEventRecord(c->ev_syrk0, c->stream));
  if(use_ws)
    k_syrk_ws<<<G, WTHREADS, WSMEM_BYTES, c->stream>>>(rowptr_dev, M, K, d, S.d_segs, S.d_cta_seg_begin, (double*)c->ws);
  else if(aligned16 && d_aligned && ((reinterpret_cast<uintptr_t>(rowptr_dev) & 15u) == 0))
    k_syrk_diag<true><<<G, THREADS, SMEM_BYTES, c->stream>>>(rowptr_dev, M, K, d, S.d_segs, S.d_cta_seg_begin, (double*)c->ws);
  else
    k_syrk_diag<false><<<G, THREADS, SMEM_BYTES, c->stream>>>(rowptr_dev, M, K, d, S.d_segs, S.d_cta_seg_begin, (double*)c->ws);
  HB_LAUNCHED();
  if(c->timing) {
    HB_CUDA(cudaEventRecord(c->ev_syrk1, c->stream));
    c->syrk_timed = true;
  }
  k_syrk_fixup<<<dim3(S.ntiles, BM / 16), 256, 0, c->stream>>>(M, S.d_tile_ij, S.d_tile_slot_begin, S.d_tile_slots, (const double*)c->ws, C, ldc);
  HB_LAUNCHED();
  return HB_OK;
}
