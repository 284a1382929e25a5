// Kernel template of the tcgen05 position GEMM (see tc_gemm.cuh).  Included where the epilogue functors live.
//
// Epilogue functor contract (one thread = one output row = one TMEM lane):
//   __device__ void chunk(long long p, long long n, long long rem, bool valid, int c0, float v[32]);
//        32 consecutive output columns [c0, c0+32) of row p (valid == p < P); called by every lane so the
//        functor may use warp collectives (the BN-statistics reduction does).
//   __device__ void finish();     called once per epilogue warp after its last tile (flush statistics)
#pragma once
#include "tc.cuh"
#include "tc_gemm.cuh"

namespace gwn {

constexpr int PGT_PRODUCERS = 128;   // warps 0-3
constexpr int PGT_MMA_WARP = 4;
constexpr int PGT_EPI_WARPS = 8;     // warps 5-12, two per TMEM lane quadrant (they split the 32-column chunks)
constexpr int PGT_THREADS = 32 * (5 + PGT_EPI_WARPS);

// lane c of the warp ends up with sum over the 32 lanes of v[c]  (31 shuffles; v is destroyed)
__device__ __forceinline__ float warp_column_sums(float v[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float keep = up ? v[i + step] : v[i];
      const float send = up ? v[i] : v[i + step];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}

template <typename Epi>
__global__ void __launch_bounds__(PGT_THREADS, 1) pos_gemm_tc_kernel(const __grid_constant__ PgParams p, Epi epi,
                                                                      int stages) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, K8 = p.n_chunks * 4;                     // 16-byte K pieces per row
  const uint32_t w_bytes = (uint32_t)K8 * (uint32_t)N * 16u;
  const uint32_t a_bytes = (uint32_t)K8 * 128u * 16u;         // one stage
  uint8_t* w_s = smem;
  uint8_t* a_s = smem + ((w_bytes + 127u) & ~127u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_s + (size_t)stages * a_bytes);
  uint64_t* full = bars;               // [stages] (<= 4)
  uint64_t* empty = bars + 4;
  uint64_t* tfull = bars + 8;          // [2]
  uint64_t* tempty = bars + 10;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  if (tid == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], PGT_PRODUCERS); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 32 * PGT_EPI_WARPS); }
    fence_barrier_init();
  }
  const uint32_t acc_cols = N <= 32 ? 32u : N <= 64 ? 64u : N <= 128 ? 128u : 256u;
  if (warp == PGT_MMA_WARP) tmem_alloc(tmem_slot, 2 * acc_cols);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w_img);
    uint4* dst = reinterpret_cast<uint4*>(w_s);
    for (int i = tid; i < (int)(w_bytes / 16); i += PGT_THREADS) dst[i] = __ldg(src + i);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PGT_MMA_WARP) {
    // ===================== producers: one row per thread =====================
    const uint32_t ro32 = (uint32_t)p.rows_per_n_out;
    int g = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int stage = g % stages, phase = (g / stages) & 1;
      mbar_wait(&empty[stage], (uint32_t)(phase ^ 1));
      const uint32_t sa = smem_u32(a_s + (size_t)stage * a_bytes) + (uint32_t)tid * 16u;
      // 32-bit row arithmetic (launcher guarantees < 2^31 positions); one IMAD.WIDE per source row
      const uint32_t pp = (uint32_t)tile * 128u + (uint32_t)tid;
      const bool pv = pp < (uint32_t)p.P;
      const uint32_t n = pp / ro32, rem = pp - n * ro32;
#pragma unroll 4
      for (int q = 0; q < p.n_chunks; ++q) {
        const int sr = (int)rem + (int)p.ch[q].row_off;
        const bool ok = pv && sr >= 0 && sr < (int)p.ch[q].rows_per_n;
        const uint32_t row = n * (uint32_t)p.ch[q].rows_per_n + (uint32_t)sr;
        const bf16* src = p.ch[q].base + (ok ? (size_t)row * (uint32_t)p.ch[q].pitch + p.ch[q].col_off : 0);
        const uint32_t d = sa + (uint32_t)(q * 4 * 128) * 16u;
        const uint32_t nb = ok ? 16u : 0u;
        cp_async16(d, src, nb);
        cp_async16(d + 2048u, src + 8, nb);
        cp_async16(d + 4096u, src + 16, nb);
        cp_async16(d + 6144u, src + 24, nb);
      }
      cp_async_commit();
      // signal the stage whose group is now guaranteed complete (stages-1 groups may stay in flight)
      if (g >= stages - 1) {
        if (stages == 2) cp_async_wait<1>(); else if (stages == 3) cp_async_wait<2>(); else cp_async_wait<3>();
        fence_proxy_async();
        mbar_arrive(&full[(g - (stages - 1)) % stages]);
      }
      ++g;
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int k = (g >= stages - 1 ? g - (stages - 1) : 0); k < g; ++k) mbar_arrive(&full[k % stages]);
  } else if (warp == PGT_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, N, false, false);
      const uint32_t w_addr = smem_u32(w_s);
      int g = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int stage = g % stages, acc = g & 1;
        mbar_wait(&tempty[acc], (uint32_t)(((g >> 1) & 1) ^ 1));
        mbar_wait(&full[stage], (uint32_t)((g / stages) & 1));
        tc_fence_after();
        const uint32_t sa = smem_u32(a_s + (size_t)stage * a_bytes);
        for (int ks = 0; ks < K8 / 2; ++ks) {
          const uint64_t adesc = make_smem_desc(sa + (uint32_t)(2 * ks) * 2048u, 2048u, 128u);
          const uint64_t bdesc = make_smem_desc(w_addr + (uint32_t)(2 * ks) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
          umma_bf16(tmem_base + (uint32_t)acc * acc_cols, adesc, bdesc, idesc, ks == 0 ? 0u : 1u);
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
        ++g;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int half = (warp - (PGT_MMA_WARP + 1)) >> 2;
    int g = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int acc = g & 1;
      mbar_wait(&tfull[acc], (uint32_t)((g >> 1) & 1));
      tc_fence_after();
      const long long pp = (long long)tile * 128 + quad * 32 + lane;
      const bool pv = pp < p.P;
      long long n = 0, rem = 0;
      if (pv) split_pos(pp, p.rows_per_n_out, n, rem);
      for (int c0 = half * 32; c0 < N; c0 += 64) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)c0, v);
        epi.chunk(pp, n, rem, pv, c0, v);
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      ++g;
    }
    epi.finish();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PGT_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * acc_cols);
  }
}

template <typename Epi>
int launch_pos_gemm_tc(PgParams& p, const Epi& epi, cudaStream_t st) {
  if (p.P <= 0) return 0;
  GWN_REQUIRE(p.P < (1ll << 31), "pos_gemm_tc: too many positions");
  GWN_REQUIRE(p.n_chunks >= 1 && p.n_chunks <= PG_TC_MAX_CHUNKS && p.N % 16 == 0 && p.N >= 16 && p.N <= 256,
              "pos_gemm_tc: unsupported shape (chunks=%d, N=%d)", p.n_chunks, p.N);
  p.n_tiles = (int)cdiv(p.P, 128);
  const size_t w_bytes = ((size_t)p.n_chunks * 4 * p.N * 16 + 127) & ~(size_t)127;
  const size_t a_bytes = (size_t)p.n_chunks * 4 * 128 * 16;
  int stages = (int)((220 * 1024 - w_bytes - 256) / a_bytes);
  if (stages > 4) stages = 4;
  GWN_REQUIRE(stages >= 2, "pos_gemm_tc: K=%d does not fit 2 stages in shared memory", 32 * p.n_chunks);
  const size_t smem = w_bytes + stages * a_bytes + 256;
  static bool attr_set = false;   // per (Epi) instantiation
  if (!attr_set) {
    GWN_CUDA(cudaFuncSetAttribute(pos_gemm_tc_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  int dev = 0, sms = 0;
  GWN_CUDA(cudaGetDevice(&dev));
  GWN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = p.n_tiles < sms ? p.n_tiles : sms;
  pos_gemm_tc_kernel<Epi><<<grid, PGT_THREADS, smem, st>>>(p, epi, stages);
  GWN_LAUNCHED();
  return 0;
}

}  // namespace gwn
