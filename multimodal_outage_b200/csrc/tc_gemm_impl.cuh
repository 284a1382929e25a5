// Kernel template of the tcgen05 position GEMM (see tc_gemm.cuh).  Included where the epilogue functors live.
//
// Epilogue functor contract (one thread = one output row = one TMEM lane):
//   __device__ void chunk(long long p, long long n, long long rem, bool valid, int c0, float v[32]);
//        32 consecutive output columns [c0, c0+32) of row p (valid == p < P); called by every lane so the
//        functor may use warp collectives (the BN-statistics reduction does).
//   __device__ void finish();     called once per epilogue warp after its last tile (flush statistics)
#pragma once
#include <cstdlib>

#include "tc.cuh"
#include "tc_gemm.cuh"
#include "tma_gemm.cuh"

namespace gwn {

constexpr int PGT_PRODUCERS = 128;   // warps 0-3 (warp 0 lane 0 issues the TMA copies)
constexpr int PGT_MMA_WARP = 4;
constexpr int PGT_EPI_WARPS = 12;    // warps 5-16, three per TMEM lane quadrant: they share the (sub-tile, 32-column chunk) items
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

struct PgMaps { CUtensorMap m[PG_TC_MAX_MAPS]; };
#define PG_TRACE(slot) do { if (p.trace && blockIdx.x == 0 && g < 48 && lane == 0) p.trace[g * 8 + (slot)] = clock64(); } while (0)

// A operand: per chunk one TMA box -> [128 rows][64 B] 64B-swizzled (K-major SW64: 8-row atoms of 512 B, the two
// K=16 halves of a chunk 32 B apart); B operand: resident no-swizzle weight image; D: two TMEM accumulators.
template <typename Epi>
__global__ void __launch_bounds__(PGT_THREADS, 1) pos_gemm_tc_kernel(const __grid_constant__ PgMaps maps,
                                                                      const __grid_constant__ PgParams p, Epi epi,
                                                                      int stages) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, K8 = p.n_chunks * 4;                     // 16-byte K pieces per weight row
  const int SUB = p.sub;
  const int NB = p.n_chunks + p.n_extra;                       // TMA boxes per sub-tile (MMA chunks + epilogue extras)
  const uint32_t a_bytes = (uint32_t)NB * 8192u * (uint32_t)SUB;   // one stage: SUB x NB boxes of 128 x 64 B
  const uint32_t w_bytes = ((uint32_t)K8 * (uint32_t)N * 16u + 1023u) & ~1023u;
  uint8_t* a_s = smem;                                        // stages first (1024-aligned boxes)
  uint8_t* w_s = smem + (size_t)stages * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_s + w_bytes);
  uint64_t* full = bars;               // [stages] (<= 4)
  uint64_t* empty = bars + 4;
  uint64_t* tfull = bars + 8;          // [2]
  uint64_t* tempty = bars + 10;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  if (tid == 0) {
    // a stage is free when the MMAs have read it and (with extras) every epilogue thread is done with it
    for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], p.n_extra ? 1 + 32 * PGT_EPI_WARPS : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 32 * PGT_EPI_WARPS); }
    fence_barrier_init();
  }
  const uint32_t acc_cols = N <= 32 ? 32u : N <= 64 ? 64u : N <= 128 ? 128u : 256u;
  const uint32_t buf_cols = acc_cols * (uint32_t)SUB;          // one accumulator buffer: SUB sub-tiles side by side
  if (warp == PGT_MMA_WARP) tmem_alloc(tmem_slot, 2 * buf_cols);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.w_img);
    uint4* dst = reinterpret_cast<uint4*>(w_s);
    for (int i = tid; i < K8 * N; i += PGT_THREADS) dst[i] = __ldg(src + i);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int g = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++g) {
        const int stage = g % stages, phase = (g / stages) & 1;
        mbar_wait(&empty[stage], (uint32_t)(phase ^ 1));
        PG_TRACE(0);
        tg::mbar_expect_tx(&full[stage], a_bytes);
        const uint32_t sa = base + (uint32_t)stage * a_bytes;
        for (int t = 0; t < SUB; ++t) {
          // sub-tile st = 128 consecutive rows of ONE sample; past the last sub-tile the sample index is out of range
          // and TMA zero-fills the box
          const int st = tile * SUB + t;
          const int n = st / p.tiles_per_n, r0 = (st - n * p.tiles_per_n) * 128;
          for (int q = 0; q < NB; ++q)
            tg::tma_3d(sa + (uint32_t)(t * NB + q) * 8192u, &maps.m[p.map_of[q]], 0, r0 + p.row_off[q], n, &full[stage]);
        }
        PG_TRACE(1);
      }
    }
    __syncwarp();
  } else if (warp == PGT_MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, N, false, false);
      const uint32_t w_addr = smem_u32(w_s);
      // descriptor templates: only the 14-bit start-address field changes per MMA (addresses < 256 KB: no carry)
      const uint64_t at = tg::make_desc_sw(0, 16u, 512u, 4u);
      const uint64_t bt = make_smem_desc(w_addr, (uint32_t)N * 16u, 128u);
      const uint32_t bstep = ((uint32_t)N * 32u) >> 4;            // two K pieces of the weight image per K=16 step
      int g = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int stage = g % stages, acc = g & 1;
        mbar_wait(&tempty[acc], (uint32_t)(((g >> 1) & 1) ^ 1));
        PG_TRACE(2);
        mbar_wait(&full[stage], (uint32_t)((g / stages) & 1));
        PG_TRACE(3);
        tc_fence_after();
        for (int t = 0; t < SUB; ++t) {
          uint64_t ad = at + (uint64_t)((base + (uint32_t)stage * a_bytes + (uint32_t)(t * NB) * 8192u) >> 4);
          uint64_t bd = bt;
          const uint32_t d = tmem_base + (uint32_t)acc * buf_cols + (uint32_t)t * acc_cols;
#pragma unroll 2
          for (int q = 0; q < p.n_chunks; ++q) {
            umma_bf16(d, ad, bd, idesc, q == 0 ? 0u : 1u);
            umma_bf16(d, ad + 2u, bd + bstep, idesc, 1u);            // second K=16 half of the chunk: +32 B
            ad += 8192u >> 4;
            bd += 2u * bstep;
          }
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
        PG_TRACE(4);
        ++g;
      }
    }
    __syncwarp();
  } else if (warp > PGT_MMA_WARP) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int rank = (warp - (PGT_MMA_WARP + 1)) >> 2;          // 0..2: which of the quadrant's three warps
    const int n_c32 = (N + 31) / 32;                            // 32-column chunks per sub-tile
    int g = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int acc = g & 1, stage = g % stages;
      mbar_wait(&tfull[acc], (uint32_t)((g >> 1) & 1));
      if (warp == PGT_MMA_WARP + 1) PG_TRACE(5);
      tc_fence_after();
      // work items (sub-tile t, chunk c) dealt round-robin to the quadrant's warps
      for (int item = rank; item < SUB * n_c32; item += PGT_EPI_WARPS / 4) {
        const int t = item / n_c32, c0 = (item - t * n_c32) * 32;
        const int st = tile * SUB + t;
        const int ns = st / p.tiles_per_n;
        const int r = (st - ns * p.tiles_per_n) * 128 + quad * 32 + lane;
        const bool pv = r < p.rows_out && ns < p.n_samples;
        // (sample, row) of the caller's view: virtual samples (position-wise GEMMs tile the flat position axis)
        long long pp = (long long)ns * p.rows_out + r, n = ns, rem = r;
        if (p.rows_out != (int)p.rows_per_n_out && pv) split_pos(pp, p.rows_per_n_out, n, rem);
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * buf_cols + (uint32_t)t * acc_cols + (uint32_t)c0, v);
        if constexpr (Epi::kExtra)
          epi.chunk_ex(pp, n, rem, pv, c0, v, smem + (size_t)stage * a_bytes + (size_t)(t * NB + p.n_chunks) * 8192, quad * 32 + lane);
        else
          epi.chunk(pp, n, rem, pv, c0, v);
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (p.n_extra) mbar_arrive(&empty[stage]);
      if (warp == PGT_MMA_WARP + 1) PG_TRACE(6);
      ++g;
    }
    epi.finish();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == PGT_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * buf_cols);
  }
}

template <typename Epi>
int launch_pos_gemm_tc(PgParams& p, const Epi& epi, cudaStream_t st) {
  if (p.P <= 0) return 0;
  GWN_REQUIRE(p.P < (1ll << 31), "pos_gemm_tc: too many positions");
  GWN_REQUIRE(p.n_chunks >= 1 && p.n_chunks <= PG_TC_MAX_CHUNKS && p.N % 16 == 0 && p.N >= 16 && p.N <= 256,
              "pos_gemm_tc: unsupported shape (chunks=%d, N=%d)", p.n_chunks, p.N);
  // position-wise GEMM (every chunk has the output's own row structure): tile the flat position axis
  bool flat = true;
  const int NB = p.n_chunks + p.n_extra;
  GWN_REQUIRE(NB <= PG_TC_MAX_CHUNKS, "pos_gemm_tc: too many chunks");
  for (int q = 0; q < NB; ++q)
    if (p.ch[q].rows_per_n != p.rows_per_n_out || p.ch[q].row_off != 0) flat = false;
  const long long n_real = p.P / p.rows_per_n_out;
  GWN_REQUIRE(n_real * p.rows_per_n_out == p.P, "pos_gemm_tc: P is not a whole number of samples");
  p.rows_out = flat ? (int)p.P : (int)p.rows_per_n_out;
  p.n_samples = flat ? 1 : (int)n_real;
  // macro tiles: several 128-row sub-tiles per pipeline step when the accumulators are narrow (the producer / MMA
  // threads pay ~0.5 us of hand-off latency per step whatever the tile holds)
  const int acc_c = p.N <= 32 ? 32 : p.N <= 64 ? 64 : p.N <= 128 ? 128 : 256;
  p.sub = 1;
  p.tiles_per_n = (int)cdiv(p.rows_out, 128);                     // 128-row sub-tiles per (virtual) sample
  const long long sub_tiles = (long long)p.n_samples * p.tiles_per_n;
  while (p.sub < 4 && 2 * acc_c * (p.sub * 2) <= 512 && (size_t)NB * 8192 * (p.sub * 2) * 2 <= 150 * 1024 &&
         cdiv(sub_tiles, p.sub * 2) >= 2 * tg_sm_count())
    p.sub *= 2;
  p.n_tiles = (int)cdiv(sub_tiles, p.sub);
  PgMaps maps;
  int n_maps = 0;
  struct Key { const bf16* base; long long rows; int pitch, col; } keys[PG_TC_MAX_MAPS];
  for (int q = 0; q < NB; ++q) {
    const PgChunk& c = p.ch[q];
    const long long rows = flat ? p.P : c.rows_per_n;
    int m = -1;
    for (int i = 0; i < n_maps; ++i)
      if (keys[i].base == c.base && keys[i].rows == rows && keys[i].pitch == c.pitch && keys[i].col == c.col_off) m = i;
    if (m < 0) {
      GWN_REQUIRE(n_maps < PG_TC_MAX_MAPS, "pos_gemm_tc: too many distinct sources");
      m = n_maps++;
      keys[m] = Key{c.base, rows, c.pitch, c.col_off};
      if (int rc = tg_map_rows3d(&maps.m[m], c.base + c.col_off, (uint64_t)rows, (uint64_t)(flat ? 1 : n_real),
                                 (uint64_t)c.pitch, 128))
        return rc;
    }
    p.map_of[q] = m;
    GWN_REQUIRE(c.row_off > -(1ll << 30) && c.row_off < (1ll << 30), "pos_gemm_tc: row offset out of range");
    p.row_off[q] = (int)c.row_off;
  }
  for (int i = n_maps; i < PG_TC_MAX_MAPS; ++i) maps.m[i] = maps.m[0];
  {
    const char* e = getenv("GWN_PG_TRACE");
    p.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 0)) : nullptr;
  }
  const size_t w_bytes = ((size_t)p.n_chunks * 4 * p.N * 16 + 1023) & ~(size_t)1023;
  const size_t a_bytes = (size_t)NB * 8192 * p.sub;
  int stages = (int)((220 * 1024 - w_bytes - 1024 - 256) / a_bytes);
  if (stages > 4) stages = 4;
  GWN_REQUIRE(stages >= 2, "pos_gemm_tc: K=%d does not fit 2 stages in shared memory", 32 * p.n_chunks);
  const size_t smem = w_bytes + stages * a_bytes + 1024 + 256;
  static bool attr_set = false;   // per (Epi) instantiation
  if (!attr_set) {
    GWN_CUDA(cudaFuncSetAttribute(pos_gemm_tc_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int sms = tg_sm_count();
  int grid = p.n_tiles < sms ? p.n_tiles : sms;
  pos_gemm_tc_kernel<Epi><<<grid, PGT_THREADS, smem, st>>>(maps, p, epi, stages);
  GWN_LAUNCHED();
  return 0;
}

}  // namespace gwn
