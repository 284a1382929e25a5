// General TMA-fed tcgen05 GEMM for the shapes that do not fit on chip (3,100-node supports, the head).
//
//   D[m, n] = sum_k A(m, k) * B(n, k)        M-tile 128, N-tile bn (<= 256), K-block 64, bf16 x bf16 -> fp32
//
// Operands are staged by TMA (cp.async.bulk.tensor) into 128B/64B-swizzled shared-memory tiles, 4-stage
// mbarrier ring; one elected thread issues tcgen05.mma (kind::f16, cta_group::1) into one of two TMEM
// accumulators (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1; persistent CTAs
// walk tiles in m-fastest order (CTAs running together share the B tile through L2).
//
// Operand staging modes (what the global tensor looks like -> UMMA canonical layout used):
//   TG_K_SW128    [MN][K] row-major, K contiguous     -> K-major,  SWIZZLE_128B, box (64 k, rows)
//   TG_MN_SW128   [K][MN] row-major, MN contiguous    -> MN-major, SWIZZLE_128B, boxes (64 mn, 64 k)
//   TG_MN_SW64    channels-last activations [slab][K=node][32 ch]: MN = (slab, ch) -> MN-major, SWIZZLE_64B,
//                 one 3-D box (32 ch, 64 nodes, bn/32 slabs); each slab is one 64-byte swizzle atom column.
//   TG_K_SW64     the same activations with K = (slab, ch) and MN = node (support gradient dA = X^T G):
//                 K-major, SWIZZLE_64B, one 3-D box (32 ch, rows nodes, 2 slabs) per 64-wide K block.
// Canonical layouts (units of 16 B), from the PTX ISA / CuTe UMMA descriptor tables:
//   K-major  SW128: ((8,m),(T,2)):((8T,SBO),(1,T))      -> SBO = 1024 B, K=16 step = +32 B
//   MN-major SW128: ((8,n),(8,k)):((1,LBO),(8,SBO))     -> LBO = next 64-wide MN atom, SBO = next 8 k-rows (1024 B)
//   MN-major SW64 : ((4,n),(8,k)):((1,LBO),(4,SBO))     -> LBO = next slab, SBO = 512 B
//   K-major  SW64 : ((8,m),(T,2)):((4T,SBO),(1,T))      -> SBO = 512 B, K=16 step = +32 B, next slab = +rows*64 B
#pragma once
#include <type_traits>
#include <cuda.h>

#include "tc.cuh"

namespace gwn {

enum { TG_K_SW128 = 0, TG_MN_SW128 = 1, TG_MN_SW64 = 2, TG_K_SW64 = 3 };

constexpr int TG_BM = 128, TG_BK = 64;
// Epilogue warps: TG_EPI_WARPS / 4 "ranks" per TMEM lane quadrant, the tile's 64-column pairs dealt round-robin to the
// ranks.  With one rank, an epilogue that loads from global memory (relu masks of the head's data gradients) exposes one
// HBM round trip per 32-column chunk: 34 us for a K = 32 GEMM whose traffic takes 11.  (GWN_TG_EPI_WARPS: A/B builds.)
#ifndef GWN_TG_EPI_WARPS
#define GWN_TG_EPI_WARPS 16
#endif
constexpr int TG_EPI_WARPS = GWN_TG_EPI_WARPS;
constexpr int TG_EPI_RANKS = TG_EPI_WARPS / 4;
static_assert(TG_EPI_WARPS % 4 == 0 && TG_EPI_WARPS >= 4 && TG_EPI_WARPS <= 16, "epilogue warps come in groups of four (TMEM lane quadrants)");
constexpr int TG_THREADS = 32 * (2 + TG_EPI_WARPS);   // warp 0 TMA, warp 1 MMA, warps 2.. epilogue

struct TgOperand {
  int mode;
  uint32_t tile_bytes;   // per stage (multiple of 1024)
  uint32_t lbo, sbo;     // descriptor byte offsets
  uint32_t koff[4];      // start-address offset of each of the 4 MMAs (K = 16) of a K block
  uint32_t layout;       // descriptor swizzle code: 2 = 128B, 4 = 64B
  int n_boxes;           // TMA boxes per stage
  uint32_t box_bytes;
  int box_mn;            // MN extent of one box
};

struct TgParams {
  int M, N, K;
  int bn;                // tile width
  int m_tiles, n_tiles, k_blocks;
  int splits;            // split-K factor (epilogue must accumulate atomically when > 1)
  int kb_per_split;
  int stages;
  int a_kwrap;           // > 0: operand A is addressed at k mod a_kwrap (split-precision GEMMs reuse A segments)
  TgOperand a, b;
};

namespace tg {
using namespace tc;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7u) << 61;
  return d;
}
__device__ __forceinline__ void load_operand(const TgOperand& o, const CUtensorMap* map, uint32_t dst, int mn0, int k0,
                                             uint64_t* bar) {
  if (o.mode == TG_K_SW128) {
    tma_2d(dst, map, k0, mn0, bar);
  } else if (o.mode == TG_MN_SW128) {
    for (int i = 0; i < o.n_boxes; ++i) tma_2d(dst + (uint32_t)i * o.box_bytes, map, mn0 + i * o.box_mn, k0, bar);
  } else if (o.mode == TG_MN_SW64) {
    tma_3d(dst, map, 0, k0, mn0 >> 5, bar);
  } else {
    tma_3d(dst, map, 0, mn0, k0 >> 5, bar);
  }
}
// ---- warp-private staging of a 32-row x 64-byte chunk (one row per lane <-> four lanes per row) ----
// The accumulator layout gives a thread one ROW, so a per-thread 16-byte global access touches 32 different rows per warp
// instruction: 32 LSU wavefronts for 512 bytes.  Through a 2 KB swizzled scratch (conflict-free in both directions) the warp
// instead moves 8 rows x 64 contiguous bytes per instruction - a quarter of the wavefronts.
constexpr int TG_WS_BYTES = 2048;
__device__ __forceinline__ uint32_t ws_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }
// q[c] = 16-byte piece c of this lane's row  ->  global rows g0 + r * ld (elements), rows [0, rows_valid)
__device__ __forceinline__ void ws_store_rows64(uint8_t* scr, const uint4 q[4], bf16* g0, long long ld, int rows_valid, int lane) {
#pragma unroll
  for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(scr + ws_off(lane, c)) = q[c];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2), c = lane & 3;
    const uint4 x = *reinterpret_cast<const uint4*>(scr + ws_off(r, c));
    if (r < rows_valid) *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(g0 + (long long)r * ld) + c * 16) = x;
  }
  __syncwarp();
}
__device__ __forceinline__ void ws_load_rows64(uint8_t* scr, uint4 q[4], const bf16* g0, long long ld, int rows_valid, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2), c = lane & 3;
    uint4 x = make_uint4(0u, 0u, 0u, 0u);
    if (r < rows_valid) x = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(g0 + (long long)r * ld) + c * 16));
    *reinterpret_cast<uint4*>(scr + ws_off(r, c)) = x;
  }
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 4; ++c) q[c] = *reinterpret_cast<const uint4*>(scr + ws_off(lane, c));
  __syncwarp();
}
__device__ __forceinline__ void pack_bf16x32(const float v[32], uint4 q[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
    __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
    __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
    q[j].x = *reinterpret_cast<uint32_t*>(&h0); q[j].y = *reinterpret_cast<uint32_t*>(&h1);
    q[j].z = *reinterpret_cast<uint32_t*>(&h2); q[j].w = *reinterpret_cast<uint32_t*>(&h3);
  }
}
}  // namespace tg

// Epilogue functors with `static constexpr bool kWarpScratch = true` are called as
//   void chunk_ws(int m0, int M, bool m_ok, int n0, float v[32], uint8_t* scratch, int lane);
// m0 = row of the warp's lane 0, scratch = the warp's private TG_WS_BYTES of shared memory (tg::ws_* helpers).
template <typename E, typename = void> struct epi_warp_scratch { static constexpr bool value = false; };
template <typename E> struct epi_warp_scratch<E, std::void_t<decltype(E::kWarpScratch)>> { static constexpr bool value = E::kWarpScratch; };

// Epilogue functor contract (one thread = one accumulator row):
//   void chunk(int m, bool m_ok, int n0, float v[32]);   32 consecutive columns [n0, n0+32) of row m
//   n0 may exceed N on the last tile: the functor masks.
template <typename Epi>
__global__ void __launch_bounds__(TG_THREADS, 1)
tma_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ TgParams p, Epi epi) {
  using namespace tc;
  using namespace tg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ST = p.stages;
  const uint32_t stage_bytes = p.a.tile_bytes + p.b.tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)ST * stage_bytes);
  uint64_t* full = bars;            // [ST] (<= 8)
  uint64_t* empty = bars + 8;       // [ST]
  uint64_t* tfull = bars + 16;      // [2]
  uint64_t* tempty = bars + 18;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  if (tid == 0) {
    for (int i = 0; i < ST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], TG_EPI_WARPS); }
    fence_barrier_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  const uint32_t acc_cols = p.bn <= 32 ? 32u : p.bn <= 64 ? 64u : p.bn <= 128 ? 128u : 256u;
  if (warp == 1) tmem_alloc(tmem_slot, 2 * acc_cols);
  pdl_wait();      // programmatic launch: barrier / TMEM set-up above overlapped the predecessor's tail (common.cuh)
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

  // producer and MMA warps: the whole warp walks the loops (addresses, coordinates and descriptors stay in uniform
  // registers, the tcgen05 / TMA instructions issue back to back); only those instructions run on one elected lane
  if (warp == 0) {
    {
      int s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile % p.m_tiles, rest = tile / p.m_tiles;
        const int nt = rest % p.n_tiles, sp = rest / p.n_tiles;
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[s], (uint32_t)(ph ^ 1));
          if (elect_one()) {
            mbar_expect_tx(&full[s], stage_bytes);
            const uint32_t sa = base + (uint32_t)s * stage_bytes, sb = sa + p.a.tile_bytes;
            load_operand(p.a, &mapA, sa, mt * TG_BM, p.a_kwrap > 0 ? (kb * TG_BK) % p.a_kwrap : kb * TG_BK, &full[s]);
            load_operand(p.b, &mapB, sb, nt * p.bn, kb * TG_BK, &full[s]);
          }
          __syncwarp();
          if (++s == ST) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {
      const uint32_t idesc = make_idesc_bf16(128, p.bn, p.a.mode == TG_MN_SW128 || p.a.mode == TG_MN_SW64,
                                             p.b.mode == TG_MN_SW128 || p.b.mode == TG_MN_SW64);
      // descriptor templates: only the 14-bit start-address field changes per MMA (addresses < 256 KB: no carry)
      const uint64_t at = make_desc_sw(0, p.a.lbo, p.a.sbo, p.a.layout);
      const uint64_t bt = make_desc_sw(0, p.b.lbo, p.b.sbo, p.b.layout);
      int s = 0, ph = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int rest = tile / p.m_tiles;
        const int sp = rest / p.n_tiles;
        const int kb0 = sp * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        const int acc = tcount & 1;
        mbar_wait(&tempty[acc], (uint32_t)(((tcount >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * acc_cols;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[s], (uint32_t)ph);
          tc_fence_after();
          const uint32_t sa = base + (uint32_t)s * stage_bytes, sb = sa + p.a.tile_bytes;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < TG_BK / 16; ++ks) {
              const uint64_t adesc = at + (uint64_t)((sa + p.a.koff[ks]) >> 4);
              const uint64_t bdesc = bt + (uint64_t)((sb + p.b.koff[ks]) >> 4);
              umma_bf16(d, adesc, bdesc, idesc, (kb == kb0 && ks == 0) ? 0u : 1u);
            }
            umma_commit(&empty[s]);
            if (kb == kb1 - 1) umma_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++s == ST) { s = 0; ph ^= 1; }
        }
        if (kb1 <= kb0) {
          if (elect_one()) umma_commit(&tfull[acc]);
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3, rank = (warp - 2) >> 2;
    int tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int mt = tile % p.m_tiles, rest = tile / p.m_tiles;
      const int nt = rest % p.n_tiles;
      const int acc = tcount & 1;
      mbar_wait(&tfull[acc], (uint32_t)((tcount >> 1) & 1));
      tc_fence_after();
      const int m = mt * TG_BM + quad * 32 + lane;
      const bool m_ok = m < p.M;
      const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * acc_cols;
      if constexpr (TG_EPI_RANKS > 1) {
        // several ranks per quadrant: one 32-column chunk at a time (the other ranks' chunks hide this one's latencies)
        for (int c0 = 32 * rank; c0 < p.bn; c0 += 32 * TG_EPI_RANKS) {
          float v[32];
          tmem_ld32(t0 + (uint32_t)c0, v);
          if constexpr (epi_warp_scratch<Epi>::value)
            epi.chunk_ws(m - lane, p.M, m_ok, nt * p.bn + c0, v, smem + (size_t)ST * stage_bytes + 256 + (size_t)(warp - 2) * TG_WS_BYTES, lane);
          else
            epi.chunk(m, m_ok, nt * p.bn + c0, v);
        }
      } else {
        for (int c0 = 0; c0 < p.bn; c0 += 64) {
          uint32_t r[2][32];
          tmem_ld32_issue(t0 + (uint32_t)c0, r[0]);
          if (c0 + 32 < p.bn) tmem_ld32_issue(t0 + (uint32_t)c0 + 32u, r[1]);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[0][j]);
          epi.chunk(m, m_ok, nt * p.bn + c0, v);
          if (c0 + 32 < p.bn) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[1][j]);
            epi.chunk(m, m_ok, nt * p.bn + c0 + 32, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);    // one arrival per warp
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * acc_cols);
  }
}

// ---- host side (tma_gemm.cu) ----
// Fills mode-dependent descriptor constants for an operand whose tile spans `rows` along MN.
void tg_operand(TgOperand& o, int mode, int rows);
// Tensor maps.  All return 0 on success.
int tg_map_2d(CUtensorMap* map, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0,
              uint32_t box1);                       // SWIZZLE_128B, bf16
int tg_map_slabs(CUtensorMap* map, const void* base, uint64_t V, uint64_t slabs, uint32_t box_v, uint32_t box_slabs);
// bf16 [rows][cols] row-major matrix with `pitch` elements per row, K-major box (64, box_rows)
inline int tg_map_rows(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows) {
  return tg_map_2d(map, base, cols, rows, pitch * 2, 64, box_rows);
}
// channels-last activation [n][rows][32 of `pitch` channels] -> 3-D map {32, rows, n}, box {32, box_rows, 1}, 64B swizzle
int tg_map_rows3d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t n, uint64_t pitch, uint32_t box_rows);
// same view, box {32, box_rows, box_n} (several samples / slabs per box)
int tg_map_rows3d_box(CUtensorMap* map, const void* base, uint64_t rows, uint64_t n, uint64_t pitch, uint32_t box_rows,
                      uint32_t box_n);
int tg_sm_count();

template <typename Epi>
int launch_tma_gemm(const CUtensorMap& mapA, const CUtensorMap& mapB, TgParams& p, const Epi& epi, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0) return 0;
  GWN_REQUIRE(p.bn % 32 == 0 && p.bn >= 32 && p.bn <= 256, "tma_gemm: bad tile width %d", p.bn);
  p.m_tiles = (int)cdiv(p.M, TG_BM);
  p.n_tiles = (int)cdiv(p.N, p.bn);
  p.k_blocks = (int)cdiv(p.K, TG_BK);
  if (p.splits < 1) p.splits = 1;
  p.kb_per_split = (int)cdiv(p.k_blocks, p.splits);
  p.splits = (int)cdiv(p.k_blocks, p.kb_per_split);
  const size_t stage = (size_t)p.a.tile_bytes + p.b.tile_bytes;
  constexpr bool kWS = epi_warp_scratch<Epi>::value;
  static_assert(!kWS || TG_EPI_RANKS > 1, "warp-scratch epilogues use the one-chunk-at-a-time epilogue loop");
  const size_t ws_bytes = kWS ? (size_t)TG_EPI_WARPS * tg::TG_WS_BYTES : 0;
  int stages = (int)((size_t)((kWS ? 227 : 224) * 1024 - 1024 - 256 - ws_bytes) / stage);
  if (stages > 6) stages = 6;
  GWN_REQUIRE(stages >= 2, "tma_gemm: tile does not fit 2 stages");
  p.stages = stages;
  const size_t smem = stages * stage + 1024 + 256 + ws_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    GWN_CUDA(cudaFuncSetAttribute(tma_gemm_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int sms = tg_sm_count();
  const long long tiles = (long long)p.m_tiles * p.n_tiles * p.splits;
  const int grid = (int)(tiles < sms ? tiles : sms);
  GWN_CUDA(launch_pdl(tma_gemm_kernel<Epi>, dim3(grid), dim3(TG_THREADS), smem, st, mapA, mapB, p, epi));
  GWN_LAUNCHED();
  return 0;
}

}  // namespace gwn
